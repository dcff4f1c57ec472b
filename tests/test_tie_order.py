"""Exact distance ties as the reference leaves them (abcb200_set_tie_order 1): the host step, on the CPU.

PLS::ordered (lib/PLS/include/PLS/pls.h:58-69; lib/ranker.h:47-53 is the same construction) is an index std::sort with a strict <:
not stable, so where equal distances land is decided by libstdc++'s introsort over all N indices. abcb200_tie_order_stdsort takes the
distances and an ascending order whose ties are placed arbitrarily (the device order: ascending particle index) and returns the first
top_n entries of PLS::ordered(dist) — re-deriving them only when exact ties reach the output. Checked here against the oracle's
`ordered` (the same statement, compiled by the same libstdc++) and, where /root/reference is present, against the reference's own
template compiled unmodified (oracle/_ref). The GPU side (distances computed on the device, the chained entry point, the dice-game
stand-in) is tests/test_gpu_chain.py."""
import ctypes as C
import os

import numpy as np
import pytest

from abcsmc_b200 import _capi, api


def _device_order(dist, top_n):
    """What the CUDA ranking returns: ascending distance, exact ties by ascending particle index."""
    return np.lexsort((np.arange(dist.size), dist))[:top_n].astype(np.uint64)


CASES = [(1000, 100, 50), (5000, 5000, 300), (257, 16, 3), (4000, 500, 100000), (20000, 1000, 700), (64, 64, 1), (3, 2, 2)]


@pytest.mark.parametrize("N,top_n,levels", CASES)
def test_host_step_reproduces_std_sort(oracle, N, top_n, levels):
    rng = np.random.default_rng(N + top_n)
    dist = np.sqrt(rng.integers(0, levels, N).astype(np.float64))                 # integer-valued metrics: many exact ties
    got, changed = api.tie_order_stdsort(dist, _device_order(dist, top_n))
    want = oracle.ordered(dist)[:top_n].astype(np.uint64)
    assert np.array_equal(got, want)
    ds = np.sort(dist)
    ties_reach_output = bool(np.any(np.diff(ds[:top_n]) == 0) or (top_n < N and ds[top_n] == ds[top_n - 1]))
    assert changed == ties_reach_output


def test_unambiguous_orders_are_left_alone(oracle):
    rng = np.random.default_rng(7)
    dist = rng.random(3000)
    dev = _device_order(dist, 300)
    got, changed = api.tie_order_stdsort(dist, dev)
    assert not changed and np.array_equal(got, dev) and np.array_equal(got, oracle.ordered(dist)[:300].astype(np.uint64))
    # ties strictly beyond the cut do not matter for the first top_n entries
    dist[dev[-1]] = 2.0; dist[5] = 3.0; dist[6] = 3.0                              # the last kept one is unique; two equal ones far behind
    dev = _device_order(dist, 300)
    got, changed = api.tie_order_stdsort(dist, dev)
    assert not changed and np.array_equal(got, oracle.ordered(dist)[:300].astype(np.uint64))
    # a tie exactly across the cut does
    far = int(np.argmax(dist))
    dist[far] = dist[dev[-1]]
    got, changed = api.tie_order_stdsort(dist, _device_order(dist, 300))
    assert changed and np.array_equal(got, oracle.ordered(dist)[:300].astype(np.uint64))


def test_bad_arguments_are_refused():
    lib = _capi.lib()
    d = np.zeros(4); o = np.array([0, 1, 9], dtype=np.uint64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.abcb200_tie_order_stdsort(p(d), 4, 3, p(o)) < 0                    # index outside 0 .. N-1
    assert lib.abcb200_tie_order_stdsort(p(d), 4, 5, p(o)) < 0                    # top_n > N
    assert lib.abcb200_tie_order_stdsort(None, 4, 3, p(o)) < 0
    dn = np.array([0.5, np.nan, 0.25, 1.0]); on = np.array([2, 0, 3], dtype=np.uint64)
    assert lib.abcb200_tie_order_stdsort(p(dn), 4, 3, p(on)) < 0                  # NaN: the comparator would be inconsistent
    assert lib.abcb200_set_tie_order(None, 1) < 0


def test_host_step_against_the_reference_template():
    """The reference's own PLS::ordered (compiled unmodified into oracle/_ref) on tied data."""
    import oracle.ref as ref
    if not os.path.exists(os.path.join(ref.REFERENCE_ROOT, "lib", "PLS", "src", "pls.cpp")):
        pytest.skip("reference sources not present (GPU box)")
    ref.build()
    rng = np.random.default_rng(99)
    for N, top_n, levels in ((900, 450, 40), (6000, 600, 500)):
        dist = np.sqrt(rng.integers(0, levels, N).astype(np.float64))
        got, _ = api.tie_order_stdsort(dist, _device_order(dist, top_n))
        assert np.array_equal(got, ref.ordered(dist)[:top_n].astype(np.uint64))
