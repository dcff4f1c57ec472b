"""Host-side logic of the row-sharded weight update on CPU: world_size-2 and -3 `gloo` groups.

The exchange step (abcsmc_b200.device.sharded_weight_update) is independent of where a rank's rows are computed:
on the GPU box `local_fn` is the CUDA kernel (abcb200_weights_unnorm_dev); here the test supplies the CPU oracle as
`local_fn` so that partitioning, padding, the all-reduce of the sum of squares and the all-gather can be checked
without a GPU. The oracle is used as the checker's stand-in for the kernel, never as a product path.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from abcsmc_b200 import device as dev
from abcsmc_b200 import synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _unnorm_rows(th_new, th_old, w_old, dv):
    """numer / den for every row (vectorised restatement of src/AbcUtil.cpp:556-581, dv > 0)."""
    sig = np.sqrt(dv)
    u = (th_new[:, None, :] - th_old[None, :, :]) / sig
    dens = np.exp(-0.5 * (u ** 2).sum(axis=2) - np.sum(np.log(np.sqrt(2 * np.pi) * sig))) @ w_old
    return 1.0 / dens


def _worker(rank, world, port, n_new, n_old, P, gather, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        th_new, th_old, w_old, dv = synth.make_weight_case(n_new, n_old, P, seed=77)

        def local_fn(lo, hi, w_loc, ss):
            w = _unnorm_rows(th_new[lo:hi], th_old, w_old, dv)
            w_loc[: hi - lo] = torch.from_numpy(w)
            ss[0] = float(np.sum(w * w))

        def scale_fn(w_loc, n, ss):
            if ss.item() > 0:
                w_loc[:n] /= torch.sqrt(ss)

        out = dev.sharded_weight_update(local_fn, scale_fn, n_new, torch.device("cpu"), gather=gather)
        q.put((rank, out.numpy().copy()))
    finally:
        dist.destroy_process_group()


def _run(world, n_new, n_old, P, gather):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_new, n_old, P, gather, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


@pytest.mark.parametrize("world,n_new", [(2, 301), (3, 100), (2, 1)])
def test_sharded_weight_update_matches_single_process(oracle, world, n_new):
    n_old, P = 120, 4
    th_new, th_old, w_old, dv = synth.make_weight_case(n_new, n_old, P, seed=77)
    want = oracle.weight_predictive_prior(np.ones(n_new), th_new, th_old, w_old, dv)
    res = _run(world, n_new, n_old, P, gather=True)
    for r in range(world):
        np.testing.assert_allclose(res[r], want, rtol=1e-10)
        assert res[r].shape == (n_new,)
    # every rank ends with the same full vector, bit for bit (one all-gather, no rank-dependent arithmetic afterwards)
    for r in range(1, world):
        assert np.array_equal(res[0], res[r])


def test_sharded_weight_update_local_slices(oracle):
    world, n_new, n_old, P = 2, 257, 90, 3
    th_new, th_old, w_old, dv = synth.make_weight_case(n_new, n_old, P, seed=77)
    want = oracle.weight_predictive_prior(np.ones(n_new), th_new, th_old, w_old, dv)
    res = _run(world, n_new, n_old, P, gather=False)
    for r in range(world):
        per, lo, hi = dev.shard_bounds(n_new, world, r)
        np.testing.assert_allclose(res[r], want[lo:hi], rtol=1e-10)


def test_shard_bounds_cover_rows_exactly():
    for n in (0, 1, 7, 8, 1000, 1_000_000):
        for world in (1, 2, 3, 4, 8):
            rows = []
            for r in range(world):
                per, lo, hi = dev.shard_bounds(n, world, r)
                assert 0 <= lo <= hi <= n and hi - lo <= per
                rows += list(range(lo, hi)) if n <= 1000 else [(lo, hi)]
            if n <= 1000:
                assert rows == list(range(n))
            else:
                assert rows[0][0] == 0 and rows[-1][1] == n and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
