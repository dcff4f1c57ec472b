"""Device-resident entry points (the `_dev` half of the C ABI) on torch CUDA tensors, and the row-sharded
multi-GPU weight update. torch is used for device memory, streams and torch.distributed only.

Layout: a column-major N x K matrix (Eigen::MatrixXd) is a contiguous torch tensor of shape (K, N).
"""
import ctypes as C

import torch

from . import _capi
from . import api as _api


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _chk(t, name):
    if t.dtype != torch.float64 or not t.is_cuda or not t.is_contiguous():
        raise ValueError(f"{name}: need a contiguous float64 CUDA tensor")


def use_torch_stream(ctx):
    """Launch the library's kernels on torch's current stream (so torch.cuda.Event brackets them)."""
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)


def rank_pls(ctx, met_t, par_t, target_t, training_fraction=0.5, top_n=0, method=0, want_dist=False):
    """abcb200_rank_pls_dev. met_t (K, N), par_t (P, N), target_t (K). Returns (order int64 [top_n], dist or None,
    n_comp_used, n_comp_per_y)."""
    for t, n in ((met_t, "met"), (par_t, "par"), (target_t, "target")):
        _chk(t, n)
    K, N = met_t.shape
    P = par_t.shape[0]
    n_out = N if top_n <= 0 or top_n > N else int(top_n)
    order = torch.empty(n_out, dtype=torch.int64, device=met_t.device)
    dist = torch.empty(N, dtype=torch.float64, device=met_t.device) if want_dist else None
    used = C.c_int(0)
    ncomp = (C.c_int32 * P)()
    ctx.check(ctx._lib.abcb200_rank_pls_dev(ctx._h, _p(met_t), N, _p(par_t), N, N, K, P, _p(target_t), float(training_fraction),
                                            int(method), n_out, _p(order), _p(dist), C.cast(C.byref(used), C.c_void_p),
                                            C.cast(ncomp, C.c_void_p)))
    return order, dist, used.value, list(ncomp)


def rank_simple(ctx, met_t, target_t, top_n=0, want_dist=False):
    _chk(met_t, "met"); _chk(target_t, "target")
    K, N = met_t.shape
    n_out = N if top_n <= 0 or top_n > N else int(top_n)
    order = torch.empty(n_out, dtype=torch.int64, device=met_t.device)
    dist = torch.empty(N, dtype=torch.float64, device=met_t.device) if want_dist else None
    ctx.check(ctx._lib.abcb200_rank_simple_dev(ctx._h, _p(met_t), N, N, K, _p(target_t), n_out, _p(order), _p(dist)))
    return order, dist


def doubled_variance_gather(ctx, par_t, order):
    """Rows par[order, :] gathered on the device (AbcSmc.cpp:1045) and their doubled variance. Returns (gathered (P, n), dv (P))."""
    _chk(par_t, "par")
    P, N = par_t.shape
    n = order.numel()
    g = torch.empty((P, n), dtype=torch.float64, device=par_t.device)
    dv = torch.empty(P, dtype=torch.float64, device=par_t.device)
    ctx.check(ctx._lib.abcb200_doubled_variance_gather_dev(ctx._h, _p(par_t), N, _p(order), n, P, _p(g), _p(dv)))
    return g, dv


def weights(ctx, numer_t, th_new_t, th_old_t, w_old_t, dv_old_t, algo=0):
    """abcb200_weights_dev: L2-normalised weights for all rows of th_new_t (P, N_new)."""
    P, n_new = th_new_t.shape
    n_old = th_old_t.shape[1]
    out = torch.empty(n_new, dtype=torch.float64, device=th_new_t.device)
    ctx.check(ctx._lib.abcb200_weights_dev(ctx._h, _p(numer_t), _p(th_new_t), n_new, n_new, _p(th_old_t), n_old, n_old, _p(w_old_t),
                                           _p(dv_old_t), P, int(algo), _p(out)))
    return out


def shard_bounds(n_rows, world, rank):
    """Rows [lo, hi) of rank `rank`; every rank's slice is padded to `per` entries for the all-gather."""
    per = (n_rows + world - 1) // world
    return per, min(rank * per, n_rows), min((rank + 1) * per, n_rows)


def sharded_weight_update(local_fn, scale_fn, n_new, device, group=None, gather=True):
    """The exchange step of the row-sharded weight update written on torch.distributed, independent of where the rows
    are computed. The product path is `weights_sharded` below (the C library's NCCL group, sharded.cu); this torch form of the same
    partitioning and exchange is what the world_size-2 / -3 gloo tests drive on CPU (tests/test_sharded_gloo.py).
    local_fn(lo, hi, w_loc, ss) fills w_loc[:hi-lo] with un-normalised weights and ss[0] with their sum of squares;
    scale_fn(w_loc, n, ss) divides by sqrt(ss) when ss > 0 (Eigen normalize(), src/AbcUtil.cpp:583).
    Collectives: one all-reduce of a double, one all-gather of the slices. No other data moves."""
    import torch.distributed as dist
    G = dist.get_world_size(group)
    r = dist.get_rank(group)
    per, lo, hi = shard_bounds(n_new, G, r)
    w_loc = torch.zeros(per, dtype=torch.float64, device=device)
    ss = torch.zeros(1, dtype=torch.float64, device=device)
    if hi > lo:
        local_fn(lo, hi, w_loc, ss)
    dist.all_reduce(ss, op=dist.ReduceOp.SUM, group=group)   # NaN rows poison the sum as they poison squaredNorm()
    if hi > lo:
        scale_fn(w_loc, hi - lo, ss)
    if not gather:
        return w_loc[: hi - lo]
    full = torch.empty(per * G, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(full, w_loc, group=group)
    return full[:n_new]


class ShardGroup:
    """abcb200_group over the ranks of a torch.distributed process group (one process per GPU): the NCCL communicator lives in
    the C library (sharded.cu), torch.distributed only carries the 128-byte id from rank 0 to the others once."""

    def __init__(self, ctx, group=None):
        import torch.distributed as dist
        self.ctx = ctx
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        dev_ = torch.device("cuda", ctx.device)
        idt = torch.zeros(128, dtype=torch.uint8, device=dev_)
        if self.rank == 0:
            buf = (C.c_uint8 * 128)()
            rc = ctx._lib.abcb200_group_unique_id(C.cast(buf, C.c_void_p), 128)
            if rc != 0:
                raise _capi.Abcb200Error(rc, "abcb200_group_unique_id failed (libnccl.so.2 not loadable?)")
            idt.copy_(torch.frombuffer(bytearray(buf), dtype=torch.uint8))
        if self.world > 1:
            dist.broadcast(idt, 0, group=group)
        raw = bytes(idt.cpu().numpy().tobytes())
        h = C.c_void_p()
        ctx.check(ctx._lib.abcb200_group_create_rank(ctx._h, C.c_char_p(raw), self.rank, self.world, C.byref(h)))
        self._h = h

    def check(self, rc):
        if rc != 0:
            raise _capi.Abcb200Error(rc, self.ctx._lib.abcb200_group_last_error(self._h).decode())

    def close(self):
        if self._h:
            self.ctx._lib.abcb200_group_destroy(self._h)
            self._h = None


def weights_sharded(ctx, numer_t, th_new_t, th_old_t, w_old_t, dv_old_t, group=None, algo=0, gather=True, shard_group=None, bcast_root=-1):
    """Row-sharded weight update over the ranks (one process per GPU): abcb200_weights_sharded_dev. Every rank holds the full
    th_new_t (P, N_new); rank r evaluates rows [r per, (r+1) per). bcast_root >= 0: the previous set (th_old_t, w_old_t, dv_old_t)
    is broadcast from that rank by the library (ncclBroadcast) before use. `shard_group`: a ShardGroup to re-use (creating one
    is a collective: every rank must do it); without one, a group is created per call.
    Returns the gathered weights (N_new) on every rank, or this rank's slice when gather is False."""
    use_torch_stream(ctx)
    P, n_new = th_new_t.shape
    n_old = th_old_t.shape[1]
    sg = shard_group or ShardGroup(ctx, group)
    try:
        per, lo, hi = shard_bounds(n_new, sg.world, sg.rank)
        full = torch.empty(n_new, dtype=torch.float64, device=th_new_t.device) if gather else None
        loc = torch.empty(per, dtype=torch.float64, device=th_new_t.device) if not gather else None
        sg.check(ctx._lib.abcb200_weights_sharded_dev(sg._h, _p(numer_t), _p(th_new_t), n_new, n_new, _p(th_old_t), n_old, n_old, _p(w_old_t),
                                                      _p(dv_old_t), P, int(algo), int(bcast_root), _p(full), _p(loc)))
        return full if gather else loc[: hi - lo]
    finally:
        if shard_group is None:
            torch.cuda.current_stream().synchronize()
            sg.close()


def host_to_colmajor_tensor(a, device, pin=False):
    """numpy (N, K) -> torch (K, N) contiguous (column-major N x K) on `device`."""
    import numpy as np
    t = torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64).T))
    if pin:
        t = t.pin_memory()
    return t.to(device, non_blocking=pin)


Context = _api.Context
