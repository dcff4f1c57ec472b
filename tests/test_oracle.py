"""Pins the CPU oracle (oracle/abc_oracle.cpp): the reference's own known-answer tests, an independent
numpy/scipy formulation on the reference's toy fixtures and seeded synthetic data, and algebraic invariants.
CPU only (no GPU, no /root/reference at run time)."""
import os

import numpy as np
import pytest

import np_reference as npr
from abcsmc_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _align(a, b):
    """per-column sign alignment (eigenvector sign is arbitrary)"""
    s = np.sign(np.sum(a * b, axis=0)); s[s == 0] = 1
    return a * s


# ---- the reference's three known-answer tests -------------------------------------------------
def test_known_answer_colwise_z_scores(oracle):
    # /root/reference/tests/abcutil.cpp:11-21
    ref = np.array([[1, 1, 1], [2, 3, 4], [3, 5, 7]], dtype=float)
    stand = np.array([[-1, -1, -1], [0, 0, 0], [1, 1, 1]], dtype=float)
    assert np.sum((stand - oracle.colwise_z_scores(ref)) ** 2) < 1e-6


def test_known_answer_euclidean(oracle):
    # /root/reference/tests/abcutil.cpp:28-40
    res = oracle.euclidean(np.array([[1, 1], [3, 3]], dtype=float), np.array([1.0, 1.0]))
    assert np.linalg.norm(res - np.array([0, 2.828427])) < 1e-6


def test_known_answer_ordered(oracle):
    # /root/reference/tests/pls.cpp:15-24
    assert list(oracle.ordered([1.0, 2.0, 3.0])) == [0, 1, 2]
    assert list(oracle.ordered([2.0, 1.0, 3.0])) == [1, 0, 2]


# ---- small pieces -----------------------------------------------------------------------------
def test_normalcdf_matches_formula(oracle):
    for z in (-3.0, -1.2815, -0.1, 0.0, 0.5, 1.2815, 4.0):
        assert oracle.normalcdf(z) == pytest.approx(npr.normalcdf(z), rel=1e-15)
    assert abs(oracle.normalcdf(1.2815) - 0.9) < 3e-4   # 4-term A&S accuracy


def test_zscore_constant_column_is_nan(oracle):
    # lib/PLS/src/pls.cpp:103 divides by the unguarded stdev
    X = np.array([[1.0, 2.0], [1.0, 3.0], [1.0, 5.0]])
    Z = oracle.colwise_z_scores(X)
    assert np.all(np.isnan(Z[:, 0])) and np.all(np.isfinite(Z[:, 1]))


def test_gaussian_pdf(oracle):
    from scipy.stats import norm
    for x, s in ((0.0, 1.0), (0.3, 0.1), (-2.0, 3.0)):
        assert oracle.gsl_ran_gaussian_pdf(x, s) == pytest.approx(norm.pdf(x, scale=s), rel=1e-14)


def test_prior_likelihoods(oracle):
    assert oracle.prior_likelihood(oracle.PRIOR_UNIFORM, 0, 2, 1.0) == 0.5
    assert oracle.prior_likelihood(oracle.PRIOR_UNIFORM, 0, 2, 2.5) == 0.0
    assert oracle.prior_likelihood(oracle.PRIOR_DISCRETE_UNIFORM, 1, 6, 3.0) == pytest.approx(1 / 6)
    assert oracle.prior_likelihood(oracle.PRIOR_DISCRETE_UNIFORM, 1, 6, 3.5) == 0.0


def test_dominant_eigenvector(oracle):
    rng = np.random.default_rng(3)
    for n in (2, 5, 30, 50):
        B = rng.standard_normal((n + 5, n)) * np.linspace(2, 0.1, n)
        S = B.T @ B
        q = oracle.dominant_eigenvector_sym(S)
        ev, evec = np.linalg.eigh(S)
        v = evec[:, -1]; v = v * np.sign(v @ q)
        assert np.max(np.abs(q - v)) < 1e-12
        assert abs(np.linalg.norm(q) - 1) < 1e-14


def test_wilcoxon_vs_scipy_ranks(oracle):
    rng = np.random.default_rng(5)
    for n in (7, 100, 5001):
        e1, e2 = rng.standard_normal(n), 1.05 * rng.standard_normal(n)
        assert oracle.wilcoxon(e1, e2) == pytest.approx(npr.wilcoxon(e1, e2), rel=1e-13)


def test_doubled_variance(oracle):
    rng = np.random.default_rng(7)
    X = 3.0 + rng.standard_normal((400, 6)) * np.arange(1, 7)
    np.testing.assert_allclose(oracle.calculate_doubled_variance(X), npr.doubled_variance(X), rtol=1e-12)
    assert np.all(oracle.calculate_doubled_variance(X[:1]) == 0.0)


# ---- PLS on the reference's demo data and synthetic sets --------------------------------------
@pytest.mark.parametrize("method", [0, 1])
def test_pls_toy_fixture(oracle, method):
    d = np.load(os.path.join(GOLD, "toy_inputs.npz"))
    X = oracle.colwise_z_scores(d["toyX"]); Y = oracle.colwise_z_scores(d["toyY"])
    A = 5
    m = oracle.Model(X, Y, method, A)
    n = npr.kernel_pls(X, Y, A, gram=bool(method))
    for name in ("W", "P", "R", "Q"):
        got = getattr(m, name); want = _align(n[name], got)
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-9 * np.abs(want).max())
    np.testing.assert_allclose(m.coefficients(A), n["R"] @ n["Q"].T, rtol=0, atol=1e-9)


def test_pls_nir_octane_single_response(oracle):
    d = np.load(os.path.join(GOLD, "toy_inputs.npz"))
    X = oracle.colwise_z_scores(d["nir"]); Y = oracle.colwise_z_scores(d["octane"])
    m = oracle.Model(X, Y, 0, 6)       # M == 1 path (pls.cpp:403-404)
    n = npr.kernel_pls(X, Y, 6)
    np.testing.assert_allclose(m.coefficients(6), n["R"] @ n["Q"].T, rtol=0, atol=1e-10)
    np.testing.assert_allclose(m.T, _align(n["T"], m.T), rtol=0, atol=1e-9)
    ev = m.explained_variance(X, Y, 6)
    assert 0.9 < ev[0] <= 1.0


def test_pls_invariants_synthetic(oracle):
    par, met, _ = synth.make_set(600, 4, 9, seed=11)
    X = oracle.colwise_z_scores(met); Y = oracle.colwise_z_scores(par)
    m1 = oracle.Model(X, Y, 0); m2 = oracle.Model(X, Y, 1)
    # full-rank PLS == OLS
    ols = np.linalg.lstsq(X, Y, rcond=None)[0]
    np.testing.assert_allclose(m1.coefficients(), ols, atol=1e-10)
    np.testing.assert_allclose(m2.coefficients(), ols, atol=1e-10)
    # T columns mutually orthogonal
    T = m1.T; G = T.T @ T
    assert np.max(np.abs(G - np.diag(np.diag(G)))) < 1e-9 * np.max(np.diag(G))
    # scores(X) reproduces T
    np.testing.assert_allclose(m1.scores(X), T, atol=1e-11)
    # residuals / SSE consistency
    E = m1.residuals(X, Y, 3)
    np.testing.assert_allclose(m1.SSE(X, Y, 3), (E ** 2).sum(axis=0), rtol=1e-13)


def test_cv_new_data_and_selection(oracle):
    par, met, _ = synth.make_set(900, 3, 8, seed=21)
    X = oracle.colwise_z_scores(met); Y = oracle.colwise_z_scores(par)
    m = oracle.Model(X[:450], Y[:450])
    res = m.cv_NEW_DATA(X[450:], Y[450:])
    n = npr.kernel_pls(X[:450], Y[:450], 8)
    E = npr.holdout_errors(n, X[450:], Y[450:])
    cube = np.stack([e for e in res.errors()])          # [y][n, c]
    np.testing.assert_allclose(cube, E, atol=1e-11)
    nc, press = npr.optimal_num_components(E)
    np.testing.assert_allclose(res.validation(oracle.RESS), press, rtol=1e-11)
    np.testing.assert_allclose(res.validation(oracle.MSE), press / 450, rtol=1e-11)
    assert list(res.optimal_num_components()) == list(nc)


def test_cv_loo_small(oracle):
    par, met, _ = synth.make_set(40, 2, 4, seed=31)
    X = oracle.colwise_z_scores(met); Y = oracle.colwise_z_scores(par)
    m = oracle.Model(X, Y, 0, 3)
    cube = np.stack(m.cv_LOO().errors())
    for i in (0, 7, 39):
        keep = np.arange(40) != i
        n = npr.kernel_pls(X[keep], Y[keep], 4)
        for c in (1, 3):
            pred = X[i] @ (n["R"][:, :c] @ n["Q"][:, :c].T)
            np.testing.assert_allclose(cube[:, i, c - 1], Y[i] - pred, atol=1e-10)


def test_cv_lso_small(oracle):
    """Model::cv_LSO (pls.cpp:512-549) against independent numpy refits on the same splits"""
    par, met, _ = synth.make_set(60, 2, 5, seed=33)
    X = oracle.colwise_z_scores(met); Y = oracle.colwise_z_scores(par)
    m = oracle.Model(X, Y, 0, 4)
    rng = np.random.default_rng(5)
    sh = np.stack([rng.permutation(60) for _ in range(3)]).astype(np.uint64)
    test_size = 15
    res = m.cv_LSO(sh, test_size)
    cube = np.stack(res.errors())                        # [y][rows, c]
    assert cube.shape == (2, 3 * test_size, 4)
    for rep in range(3):
        tr, te = sh[rep, :45].astype(int), sh[rep, 45:].astype(int)
        n = npr.kernel_pls(X[tr], Y[tr], 5)
        for c in (1, 2, 4):
            pred = X[te] @ (n["R"][:, :c] @ n["Q"][:, :c].T)
            np.testing.assert_allclose(cube[:, rep * test_size:(rep + 1) * test_size, c - 1], (Y[te] - pred).T, atol=1e-10)
    press = res.validation(oracle.RESS)
    np.testing.assert_allclose(press, np.sum(cube ** 2, axis=1), rtol=1e-12)


@pytest.mark.parametrize("shape", [(2000, 3, 6), (5000, 10, 20)])
def test_particle_ranking_pls_vs_numpy(oracle, shape):
    N, P, K = shape
    par, met, target = synth.make_set(N, P, K, seed=100 + K)
    o = oracle.particle_ranking_PLS(met, par, target, 0.5)
    n = npr.rank_pls(met, par, target, 0.5)
    assert o["ncomp_used"] == n["ncomp_used"]
    assert list(o["ncomp"]) == list(n["ncomp"])
    np.testing.assert_allclose(o["dist"], n["dist"], rtol=1e-10)
    np.testing.assert_allclose(o["press"], n["press"], rtol=1e-10)
    assert np.array_equal(o["order"].astype(np.int64), n["order"])
    assert np.all(np.diff(o["dist"][o["order"].astype(np.int64)]) >= 0)


def test_particle_ranking_simple(oracle):
    par, met, target = synth.make_set(3000, 3, 6, seed=41)
    o = oracle.particle_ranking_simple(met, target)
    z, mu, sd = npr.zscore_cols(met)
    d = np.sqrt((((z - (target - mu) / sd)) ** 2).sum(axis=1))
    np.testing.assert_allclose(o["dist"], d, rtol=1e-12)
    assert np.array_equal(o["order"].astype(np.int64), np.argsort(d, kind="stable"))


# ---- weights -----------------------------------------------------------------------------------
def test_weights_vs_numpy_and_invariants(oracle):
    th_new, th_old, w_old, dv = synth.make_weight_case(300, 200, 5, seed=51)
    numer = np.ones(300)
    w = oracle.weight_predictive_prior(numer, th_new, th_old, w_old, dv)
    np.testing.assert_allclose(w, npr.weights(numer, th_new, th_old, w_old, dv), rtol=1e-11)
    assert abs(np.sum(w * w) - 1.0) < 1e-13            # L2-normalised (src/AbcUtil.cpp:583)
    w2 = oracle.weight_predictive_prior(numer, th_new, th_old, 7.5 * w_old, dv)
    np.testing.assert_allclose(w, w2, rtol=1e-12)      # scale of w_old cancels
    assert np.all(oracle.weight_predictive_prior0(4) == 0.25)


def test_weights_converged_parameter(oracle):
    # dv == 0 and equal values: factor skipped (src/AbcUtil.cpp:573)
    th_new, th_old, w_old, dv = synth.make_weight_case(50, 40, 3, seed=61)
    th_new[:, 1] = 0.25; th_old[:, 1] = 0.25; dv[1] = 0.0
    w = oracle.weight_predictive_prior(np.ones(50), th_new, th_old, w_old, dv)
    keep = [0, 2]
    w_ref = oracle.weight_predictive_prior(np.ones(50), th_new[:, keep], th_old[:, keep], w_old, dv[keep])
    np.testing.assert_allclose(w, w_ref, rtol=1e-13)
    # dv == 0 but values differ: inf * 0 = NaN for that row; Eigen normalize() then leaves the vector unscaled
    th_new[3, 1] = 0.5
    w = oracle.weight_predictive_prior(np.ones(50), th_new, th_old, w_old, dv)
    assert np.isnan(w[3]) and np.all(np.isfinite(np.delete(w, 3)))
