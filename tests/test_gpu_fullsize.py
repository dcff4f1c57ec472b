"""The CUDA path at BASELINE.json's full sizes (configs[1] = C2 and configs[2] = C3), checked through size-independent
properties where the CPU oracle would take minutes (C3: ~140 s per pass), and against the oracle directly where it takes
about a second (C2). Properties: the order is a sorted top-N selection of the returned distances; the fused projection +
distance equals the PLS::Model API composed by hand (standardise -> fit on the training half -> hold-out selection -> scores
-> euclidean); runs are bitwise repeatable; weights are L2-normalised, invariant to the scale of the previous weights, equal
between the pairwise-difference and the DMMA formulation, and a subset of new rows gives proportional weights.
Run on the B200 box: python -m pytest tests -m gpu"""
import numpy as np
import pytest

from abcsmc_b200 import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-10


@pytest.fixture(scope="module")
def api():
    from abcsmc_b200 import api as a
    a.get_context(0)
    return a


@pytest.fixture(scope="module")
def c3():
    return synth.make_config("C3")


def _check_sorted_selection(order, dist, top_n):
    order = order.astype(np.int64)
    assert order.size == top_n and np.unique(order).size == top_n and order.min() >= 0 and order.max() < dist.size
    d = dist[order]
    assert np.all(d[1:] >= d[:-1])                                   # PLS::ordered: ascending (pls.h:58-69)
    ties = d[1:] == d[:-1]
    assert np.all(order[1:][ties] > order[:-1][ties])                # documented tie order: ascending particle index
    rest = np.ones(dist.size, dtype=bool); rest[order] = False
    assert dist[rest].min() >= d[-1]                                 # nothing better was left out (AbcSmc.cpp:645-646)


def test_c2_full_size_matches_oracle(api, oracle):
    """configs[1] at its full size: the oracle needs about one second here, so this is direct parity."""
    cfg = synth.make_config("C2")
    o = oracle.particle_ranking_PLS(cfg["metrics"], cfg["params"], cfg["target"], 0.5)
    g = api.particle_ranking_PLS(cfg["metrics"], cfg["params"], cfg["target"], 0.5, top_n=cfg["N_pp"], return_info=True)
    assert g["ncomp_used"] == o["ncomp_used"] and list(g["ncomp"]) == [int(v) for v in o["ncomp"]]
    np.testing.assert_allclose(g["dist"], o["dist"], rtol=RTOL)
    assert np.array_equal(g["order"].astype(np.int64), o["order"][:cfg["N_pp"]].astype(np.int64))
    _check_sorted_selection(g["order"], g["dist"], cfg["N_pp"])
    sel = cfg["params"][g["order"].astype(np.int64), :]
    np.testing.assert_allclose(api.calculate_doubled_variance(sel), oracle.calculate_doubled_variance(sel), rtol=RTOL)
    w = api.weight_predictive_prior(None, sel, cfg["theta_old"], cfg["w_old"], cfg["dv_old"])
    np.testing.assert_allclose(w, oracle.weight_predictive_prior(np.ones(len(sel)), sel, cfg["theta_old"], cfg["w_old"], cfg["dv_old"]), rtol=RTOL)


def test_c3_full_size_ranking_properties(api, c3):
    cfg = c3
    N, K, P, top = cfg["N"], cfg["K"], cfg["P"], cfg["N_pp"]
    g = api.particle_ranking_PLS(cfg["metrics"], cfg["params"], cfg["target"], 0.5, top_n=top, return_info=True)
    assert 1 <= g["ncomp_used"] <= K and g["ncomp_used"] == max(g["ncomp"]) and min(g["ncomp"]) >= 1
    assert np.all(np.isfinite(g["dist"])) and g["dist"].shape == (N,)
    _check_sorted_selection(g["order"], g["dist"], top)
    # bitwise repeatable (every reduction has a fixed order)
    g2 = api.particle_ranking_PLS(cfg["metrics"], cfg["params"], cfg["target"], 0.5, top_n=top, return_info=True)
    assert np.array_equal(g["order"], g2["order"]) and np.array_equal(g["dist"], g2["dist"]) and list(g["ncomp"]) == list(g2["ncomp"])
    # the same result composed by hand from the PLS::Model API, the way AbcUtil.cpp:432-455 is written
    mean_x, sd_x = api.colwise_moments(cfg["metrics"])
    X = api.colwise_z_scores(cfg["metrics"]); Y = api.colwise_z_scores(cfg["params"])
    n_tr = int(np.floor(N * 0.5 + 0.5))
    m = api.Model(X[:n_tr], Y[:n_tr])
    press, ncomp = m.cv_NEW_DATA(X[n_tr:], Y[n_tr:], alpha=0.1)
    assert list(ncomp) == list(g["ncomp"])
    assert np.all(press >= 0)
    used = int(max(ncomp))
    obs = ((cfg["target"] - mean_x) / sd_x).reshape(1, K)
    d = api.euclidean(m.scores(X, used), m.scores(obs, used).ravel())
    np.testing.assert_allclose(d, g["dist"], rtol=RTOL)
    # full order = the top-N order extended
    full = api.particle_ranking_PLS(cfg["metrics"], cfg["params"], cfg["target"], 0.5)
    assert np.array_equal(full[:top], g["order"])
    df = g["dist"][full.astype(np.int64)]
    assert np.all(df[1:] >= df[:-1]) and np.array_equal(np.sort(full.astype(np.int64)), np.arange(N))


def test_c3_full_size_weight_properties(api, c3):
    cfg = c3
    r = np.random.default_rng(5)
    P, n_new = cfg["P"], cfg["N_pp"]
    theta_old, w_old, dv_old = cfg["theta_old"], cfg["w_old"], cfg["dv_old"]
    theta_new = np.asfortranarray(np.clip(theta_old[r.integers(0, theta_old.shape[0], n_new)] + r.normal(0, 1, (n_new, P)) * np.sqrt(dv_old), 0, 1))
    w = api.weight_predictive_prior(None, theta_new, theta_old, w_old, dv_old)
    assert w.shape == (n_new,) and np.all(w > 0)
    np.testing.assert_allclose(np.linalg.norm(w), 1.0, rtol=1e-13)                       # Eigen normalize(), AbcUtil.cpp:583
    # any scale of the previous weights cancels in the normalisation
    np.testing.assert_allclose(api.weight_predictive_prior(None, theta_new, theta_old, 37.5 * w_old, dv_old), w, rtol=RTOL)
    # the reference's pairwise-difference formulation and the DMMA inner-product formulation agree
    w1 = api.weight_predictive_prior(None, theta_new, theta_old, w_old, dv_old, algo=1)
    w2 = api.weight_predictive_prior(None, theta_new, theta_old, w_old, dv_old, algo=2)
    np.testing.assert_allclose(w1, w2, rtol=RTOL)
    # rows are independent up to the common norm: a subset of new particles gets proportional weights
    sub = np.sort(r.choice(n_new, 777, replace=False))
    ws = api.weight_predictive_prior(None, np.asfortranarray(theta_new[sub]), theta_old, w_old, dv_old)
    np.testing.assert_allclose(ws / np.linalg.norm(ws), w[sub] / np.linalg.norm(w[sub]), rtol=RTOL)
    # a prior-likelihood numerator of zero gives weight zero, not NaN (particle outside a uniform prior's support)
    numer = np.ones(n_new); numer[::100] = 0.0
    wz = api.weight_predictive_prior(numer, theta_new, theta_old, w_old, dv_old)
    assert np.all(wz[::100] == 0.0) and np.all(np.isfinite(wz))
    # doubled variance of the gathered rows against a two-pass numpy variance
    np.testing.assert_allclose(api.calculate_doubled_variance(theta_new), 2.0 * theta_new.var(axis=0, ddof=1), rtol=1e-12)
