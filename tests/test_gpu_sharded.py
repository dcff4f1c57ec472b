"""The row-sharded weight update on hardware (SURVEY.md §8 row e; reference: src/AbcUtil.cpp:547-586).

1. Two row slices evaluated with the per-slice entry points, reduced by hand, against the unsharded call and the oracle.
2. abcb200_weights_sharded_dev through a one-member group (the NCCL path of abcsmc_b200.device.weights_sharded).
3. A C++ host without Python (tests/cpp/sharded_test.cpp): abcb200_group_create + abcb200_weights_sharded over every visible GPU.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from abcsmc_b200 import _capi, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case(n_new, n_old, P, seed=0xC4):
    return synth.make_weight_case(n_new, n_old, P, seed)


@pytest.mark.parametrize("algo", [0, 1, 2])
def test_two_slices_by_hand_match_unsharded_and_oracle(oracle, algo):
    import torch
    from abcsmc_b200 import api, device as dev
    n_new, n_old, P = 3001, 2000, 30          # odd split: 1501 + 1500 rows
    th_new, th_old, w_old, dv = _case(n_new, n_old, P)
    numer = 0.25 + synth.uniform(5, n_new, 3)
    ctx = api.get_context(0)
    d = torch.device("cuda", 0)
    dev.use_torch_stream(ctx)
    t_new = dev.host_to_colmajor_tensor(th_new, d); t_old = dev.host_to_colmajor_tensor(th_old, d)
    t_w = torch.from_numpy(w_old).to(d); t_dv = torch.from_numpy(dv).to(d); t_num = torch.from_numpy(numer).to(d)
    whole = dev.weights(ctx, t_num, t_new, t_old, t_w, t_dv, algo=algo).cpu().numpy()
    parts, sums = [], []
    for r in range(2):
        per, lo, hi = dev.shard_bounds(n_new, 2, r)
        w_loc = torch.zeros(per, dtype=torch.float64, device=d); ss = torch.zeros(1, dtype=torch.float64, device=d)
        ctx.check(ctx._lib.abcb200_weights_unnorm_dev(ctx._h, C.c_void_p(t_num.data_ptr() + 8 * lo), C.c_void_p(t_new.data_ptr() + 8 * lo), n_new, hi - lo,
                                                      C.c_void_p(t_old.data_ptr()), n_old, n_old, C.c_void_p(t_w.data_ptr()), C.c_void_p(t_dv.data_ptr()), P, algo,
                                                      C.c_void_p(w_loc.data_ptr()), C.c_void_p(ss.data_ptr())))
        parts.append((w_loc, hi - lo)); sums.append(ss)
    total = sums[0] + sums[1]                                  # the all-reduce, by hand
    for w_loc, n in parts:
        ctx.check(ctx._lib.abcb200_scale_weights_dev(ctx._h, C.c_void_p(w_loc.data_ptr()), n, C.c_void_p(total.data_ptr())))
    got = torch.cat([w[:n] for w, n in parts]).cpu().numpy()
    want = oracle.weight_predictive_prior(numer, th_new, th_old, w_old, dv)
    np.testing.assert_allclose(got, want, rtol=1e-10)
    np.testing.assert_allclose(got, whole, rtol=1e-12)


def test_sharded_dev_one_member_group(oracle):
    import torch
    from abcsmc_b200 import api, device as dev
    n_new, n_old, P = 2500, 1800, 12
    th_new, th_old, w_old, dv = _case(n_new, n_old, P, seed=0xC5)
    ctx = api.get_context(0)
    d = torch.device("cuda", 0)
    t_new = dev.host_to_colmajor_tensor(th_new, d); t_old = dev.host_to_colmajor_tensor(th_old, d)
    t_w = torch.from_numpy(w_old).to(d); t_dv = torch.from_numpy(dv).to(d)
    sg = dev.ShardGroup(ctx)
    try:
        full = dev.weights_sharded(ctx, None, t_new, t_old, t_w, t_dv, gather=True, shard_group=sg, bcast_root=0)
        loc = dev.weights_sharded(ctx, None, t_new, t_old, t_w, t_dv, gather=False, shard_group=sg)
        torch.cuda.synchronize()
    finally:
        sg.close()
    want = oracle.weight_predictive_prior(np.ones(n_new), th_new, th_old, w_old, dv)
    np.testing.assert_allclose(full.cpu().numpy(), want, rtol=1e-10)
    np.testing.assert_allclose(loc.cpu().numpy(), want, rtol=1e-10)


def test_cpp_host_sharded_over_visible_gpus(tmp_path, oracle):
    import torch
    _capi.build()
    exe = str(tmp_path / "sharded_test")
    libdir = os.path.join(ROOT, "abcsmc_b200")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wpedantic", "-Werror", "-O1", os.path.join(ROOT, "tests", "cpp", "sharded_test.cpp"),
                           f"-L{libdir}", "-labcsmc_b200", f"-Wl,-rpath,{libdir}", "-o", exe])
    n_new, n_old, P = 3001, 2000, 30
    for g in sorted({1, min(2, torch.cuda.device_count()), torch.cuda.device_count()}):
        out = tmp_path / f"sh{g}.bin"
        r = subprocess.run([exe, str(g), str(n_new), str(n_old), str(P), str(out)], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        raw = np.fromfile(out, dtype=np.float64)
        o = 0
        def take(n):
            nonlocal o
            a = raw[o:o + n]; o += n
            return a
        th_new = take(n_new * P).reshape(P, n_new).T; th_old = take(n_old * P).reshape(P, n_old).T
        w_old, dv, numer, w = take(n_old), take(P), take(n_new), take(n_new)
        np.testing.assert_allclose(w, oracle.weight_predictive_prior(numer, th_new, th_old, w_old, dv), rtol=1e-10)
