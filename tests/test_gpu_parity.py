"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Bit-exact for indices / component counts; 1e-10 relative for FP64 outputs (BASELINE.json north_star).
Run on the B200 box: python -m pytest tests -m gpu"""
import os

import numpy as np
import pytest

from abcsmc_b200 import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-10
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def api():
    from abcsmc_b200 import api as a
    a.get_context(0)
    return a


def _align(a, b):
    s = np.sign(np.sum(a * b, axis=0)); s[s == 0] = 1
    return a * s


def _assert_order_parity(order_gpu, order_cpu, dist_cpu, top_n):
    """bit-exact indices; a mismatch is tolerated only if the swapped particles' CPU distances are a
    near-tie (< 1e-12 relative), SURVEY.md §7 hard part 1."""
    og = np.asarray(order_gpu[:top_n], dtype=np.int64); oc = np.asarray(order_cpu[:top_n], dtype=np.int64)
    bad = np.nonzero(og != oc)[0]
    for i in bad:
        da, db = dist_cpu[og[i]], dist_cpu[oc[i]]
        assert abs(da - db) <= 1e-12 * max(abs(da), abs(db)), f"rank {i}: gpu {og[i]} vs cpu {oc[i]} not a near-tie"
    return len(bad)


# ---- reference known-answers through the CUDA path -----------------------------------------------------
def test_known_answers(api):
    z = api.colwise_z_scores(np.array([[1, 1, 1], [2, 3, 4], [3, 5, 7]], dtype=float))
    assert np.sum((z - np.array([[-1, -1, -1], [0, 0, 0], [1, 1, 1]])) ** 2) < 1e-6      # tests/abcutil.cpp:11-21
    d = api.euclidean(np.array([[1, 1], [3, 3]], dtype=float), np.array([1.0, 1.0]))
    assert np.linalg.norm(d - np.array([0, 2.828427])) < 1e-6                               # tests/abcutil.cpp:28-40
    assert list(api.ordered([1.0, 2.0, 3.0])) == [0, 1, 2]                                  # tests/pls.cpp:15-24
    assert list(api.ordered([2.0, 1.0, 3.0])) == [1, 0, 2]


def test_moments_zscores(api, oracle):
    par, met, _ = synth.make_set(12345, 3, 7, seed=5)
    met[:, 2] = met[:, 2] * 1e3 + 1e6          # large mean / spread ratio
    mean, sd = api.colwise_moments(met)
    np.testing.assert_allclose(mean, oracle.colwise_mean(met), rtol=1e-13)
    np.testing.assert_allclose(sd, oracle.colwise_stdev(met), rtol=1e-11)
    # column 2 has mean / sd ~ 1e3: a z-score near 0 is the difference of two numbers 1e3 times larger, so 1e-10 is asked relative to
    # the scale of z (O(1)), not element by element (measured: 9e-12 absolute, which is what the oracle's own rounding is worth there)
    np.testing.assert_allclose(api.colwise_z_scores(met), oracle.colwise_z_scores(met), rtol=1e-10, atol=1e-10)
    z = api.colwise_z_scores(np.array([[1.0, 2.0], [1.0, 3.0], [1.0, 5.0]]))
    assert np.all(np.isnan(z[:, 0])) and np.all(np.isfinite(z[:, 1]))   # pls.cpp:103 quirk kept


def test_ordered_ties_and_negatives(api, oracle):
    rng = np.random.default_rng(1)
    v = rng.standard_normal(70001)
    assert np.array_equal(api.ordered(v).astype(np.int64), np.argsort(v, kind="stable"))
    v = np.round(v[:5000], 1)                   # many exact ties: ascending index within ties
    assert np.array_equal(api.ordered(v).astype(np.int64), np.argsort(v, kind="stable"))
    assert np.array_equal(api.ordered(np.array([3.0])), np.array([0], dtype=np.uint64))
    with pytest.raises(Exception):
        api.ordered(np.array([1.0, np.nan, 0.0]))


@pytest.mark.parametrize("n,top", [(4096, 1), (100000, 1000), (250001, 5000), (70001, 12288), (50000, 12289)])
def test_ordered_top_n_select(api, n, top):
    rng = np.random.default_rng(n)
    for v in (rng.standard_normal(n), np.abs(rng.standard_normal(n)) * 1e-3 + 2.0, np.round(rng.standard_normal(n), 2),
              rng.standard_normal(n) * 10.0 ** rng.integers(-30, 30, n)):
        assert np.array_equal(api.ordered(v, top_n=top).astype(np.int64), np.argsort(v, kind="stable")[:top])
    v = np.full(n, 3.25); v[n // 2] = 1.0          # one bin holds everything: the select path must hand over to the full sort
    want = np.argsort(v, kind="stable")[:top]
    assert np.array_equal(api.ordered(v, top_n=top).astype(np.int64), want)
    with pytest.raises(Exception):
        v[7] = np.nan
        api.ordered(v, top_n=top)


def test_wilcoxon(api, oracle):
    rng = np.random.default_rng(2)
    for n in (7, 1000, 50001):
        e1, e2 = rng.standard_normal(n), 1.02 * rng.standard_normal(n)
        assert api.wilcoxon(e1, e2) == pytest.approx(oracle.wilcoxon(e1, e2), rel=1e-12, abs=1e-15)
    e1 = rng.standard_normal(100); e2 = e1.copy(); e2[:50] *= 1.5    # half the differences are exactly zero
    assert api.wilcoxon(e1, e2) == pytest.approx(oracle.wilcoxon(e1, e2), rel=1e-12)


def test_doubled_variance(api, oracle):
    rng = np.random.default_rng(3)
    X = 5.0 + rng.standard_normal((5000, 30)) * np.linspace(0.01, 3, 30)
    np.testing.assert_allclose(api.calculate_doubled_variance(X), oracle.calculate_doubled_variance(X), rtol=RTOL)
    assert np.all(api.calculate_doubled_variance(X[:1]) == 0.0)
    X[:, 4] = 0.25
    assert api.calculate_doubled_variance(X)[4] == 0.0


# ---- the Gram products plsr starts from (pls.cpp:396, :398): every tiling regime of gram.cu -----------------------
@pytest.mark.parametrize("shape", [(37, 3, 2), (1000, 9, 4), (5001, 20, 10), (4099, 64, 7), (3000, 150, 30), (777, 401, 1), (2100, 500, 50), (1500, 250, 60), (1200, 1000, 30)])
def test_gram_products(api, shape):
    n, K, M = shape
    rng = np.random.default_rng(n + K)
    X = rng.standard_normal((n, K)); Y = rng.standard_normal((n, M)) + 0.5 * X[:, :1]
    xx, xy = api.gram(X, Y)
    rx, ry = X.T @ X, X.T @ Y
    np.testing.assert_allclose(xx, rx, rtol=0, atol=1e-12 * np.abs(rx).max())
    np.testing.assert_allclose(xy, ry, rtol=0, atol=1e-12 * np.abs(ry).max())
    assert np.array_equal(xx, xx.T)                                   # mirrored, bitwise symmetric
    xx2, xy2 = api.gram(X, Y)
    assert np.array_equal(xx, xx2) and np.array_equal(xy, xy2)        # deterministic


# ---- PLS::Model -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("method", [0, 1, 2])
@pytest.mark.parametrize("shape", [(600, 4, 9), (4000, 10, 20), (3000, 30, 60), (2500, 50, 130)])
def test_pls_model(api, oracle, method, shape):
    N, P, K = shape
    par, met, _ = synth.make_set(N, P, K, seed=N + K)
    X = oracle.colwise_z_scores(met); Y = oracle.colwise_z_scores(par)
    g = api.Model(X, Y, method); o = oracle.Model(X, Y, method % 2)      # 2 = KERNEL_TYPE1 streamed: same results as 0
    scale = lambda a: np.abs(a).max()
    for name in ("W", "P", "R", "Q"):
        a, b = getattr(g, name), getattr(o, name)
        np.testing.assert_allclose(_align(a, b), b, rtol=0, atol=1e-10 * scale(b), err_msg=name)
    np.testing.assert_allclose(g.coefficients(), o.coefficients(), rtol=0, atol=RTOL * scale(o.coefficients()))
    for c in (1, K // 2):
        np.testing.assert_allclose(g.coefficients(c), o.coefficients(c), rtol=0, atol=RTOL * scale(o.coefficients(c)))
    if method != 1:
        np.testing.assert_allclose(_align(g.T, o.T), o.T, rtol=0, atol=1e-10 * scale(o.T))
    sc_g, sc_o = g.scores(X[:777], 3), o.scores(X[:777], 3)
    np.testing.assert_allclose(_align(sc_g, sc_o), sc_o, rtol=0, atol=RTOL * scale(sc_o))
    np.testing.assert_allclose(g.fitted_values(X[:500], 2), o.fitted_values(X[:500], 2), rtol=0, atol=RTOL)
    np.testing.assert_allclose(g.residuals(X[:500], Y[:500], 2), o.residuals(X[:500], Y[:500], 2), rtol=0, atol=RTOL)
    np.testing.assert_allclose(g.SSE(X, Y, 3), o.SSE(X, Y, 3), rtol=RTOL)


def test_pls_single_response_nir(api, oracle):
    d = np.load(os.path.join(GOLD, "toy_inputs.npz"))
    X = oracle.colwise_z_scores(d["nir"]); Y = oracle.colwise_z_scores(d["octane"])
    g = api.Model(X, Y, 0, 6); o = oracle.Model(X, Y, 0, 6)        # M == 1 (pls.cpp:403-404), K = 401
    np.testing.assert_allclose(g.coefficients(6), o.coefficients(6), rtol=0, atol=1e-10)
    X2 = oracle.colwise_z_scores(d["toyX"]); Y2 = oracle.colwise_z_scores(d["toyY"])
    g2 = api.Model(X2, Y2, 1, 5); o2 = oracle.Model(X2, Y2, 1, 5)
    np.testing.assert_allclose(g2.coefficients(5), o2.coefficients(5), rtol=0, atol=1e-10 * np.abs(o2.coefficients(5)).max())


def test_cv_new_data(api, oracle):
    par, met, _ = synth.make_set(3000, 5, 12, seed=77)
    X = oracle.colwise_z_scores(met); Y = oracle.colwise_z_scores(par)
    g = api.Model(X[:1500], Y[:1500]); o = oracle.Model(X[:1500], Y[:1500])
    res = o.cv_NEW_DATA(X[1500:], Y[1500:])
    press, ncomp = g.cv_NEW_DATA(X[1500:], Y[1500:])
    np.testing.assert_allclose(press, res.validation(oracle.RESS), rtol=RTOL)
    assert list(ncomp) == [int(v) for v in res.optimal_num_components()]
    mse, _ = g.cv_NEW_DATA(X[1500:], Y[1500:], out_type=api.MSE)
    np.testing.assert_allclose(mse, res.validation(oracle.MSE), rtol=RTOL)


def test_explained_variance(api, oracle):
    par, met, _ = synth.make_set(1200, 4, 9, seed=78)
    X = oracle.colwise_z_scores(met); Y = oracle.colwise_z_scores(par)
    g = api.Model(X[:800], Y[:800]); o = oracle.Model(X[:800], Y[:800])
    for comp in (1, 4, 9):
        np.testing.assert_allclose(g.explained_variance(X[800:], Y[800:], comp), o.explained_variance(X[800:], Y[800:], comp), rtol=1e-10, atol=1e-12)


def test_residual_select_from_cube(api, oracle):
    """PLS::validation / optimal_num_components on a materialised PLS::Residual (pls.cpp:235-289)"""
    par, met, _ = synth.make_set(2400, 4, 10, seed=79)
    X = oracle.colwise_z_scores(met); Y = oracle.colwise_z_scores(par)
    res = oracle.Model(X[:1200], Y[:1200]).cv_NEW_DATA(X[1200:], Y[1200:])
    cube = np.stack([e.T for e in res.errors()])                  # [y][c][i]
    r = api.Residual(cube, "NEW DATA")
    np.testing.assert_allclose(r.validation(api.RESS), res.validation(oracle.RESS), rtol=RTOL)
    np.testing.assert_allclose(r.validation(api.MSE), res.validation(oracle.MSE), rtol=RTOL)
    for alpha in (0.1, 0.5, 0.01):
        assert list(r.optimal_num_components(alpha)) == [int(v) for v in res.optimal_num_components(alpha)]


@pytest.mark.parametrize("shape,A", [((60, 2, 4), 3), ((400, 3, 8), 8), ((700, 10, 20), 20), ((300, 1, 6), 6), ((260, 5, 40), 17), ((330, 12, 150), 30)])
def test_cv_loo(api, oracle, shape, A):
    """Model::cv_LOO (pls.cpp:469-491): batched down-dated refits on chip vs the oracle's N literal refits"""
    N, P, K = shape
    par, met, _ = synth.make_set(N, P, K, seed=500 + N)
    X = oracle.colwise_z_scores(met); Y = oracle.colwise_z_scores(par)
    ro = oracle.Model(X, Y, 0, A).cv_LOO()
    rg = api.Model(X, Y, 0, A).cv_LOO()
    co = np.stack([e.T for e in ro.errors()])
    scale = np.abs(co).max()
    np.testing.assert_allclose(rg.cube, co, rtol=1e-10, atol=1e-10 * scale)
    np.testing.assert_allclose(rg.validation(api.RESS), ro.validation(oracle.RESS), rtol=1e-10)
    assert list(rg.optimal_num_components()) == [int(v) for v in ro.optimal_num_components()]


def test_cv_loo_wide_predictors_streamed(api, oracle):
    """K above the on-chip limit: down-dated Gram matrices through the L2-streamed component loop"""
    N, P, K, A = 230, 3, 200, 12
    par, met, _ = synth.make_set(N, P, K, seed=901)
    X = oracle.colwise_z_scores(met); Y = oracle.colwise_z_scores(par)
    ro = oracle.Model(X, Y, 0, A).cv_LOO()
    rg = api.Model(X, Y, 0, A).cv_LOO()
    co = np.stack([e.T for e in ro.errors()])
    np.testing.assert_allclose(rg.cube, co, rtol=1e-10, atol=1e-10 * np.abs(co).max())
    assert list(rg.optimal_num_components()) == [int(v) for v in ro.optimal_num_components()]


@pytest.mark.parametrize("method", [0, 1])
def test_cv_lso(api, oracle, method):
    """Model::cv_LSO (pls.cpp:512-549) on caller-supplied splits"""
    N, P, K, A, trials, test_size = 900, 4, 12, 12, 4, 225
    par, met, _ = synth.make_set(N, P, K, seed=601)
    X = oracle.colwise_z_scores(met); Y = oracle.colwise_z_scores(par)
    rng = np.random.default_rng(7)
    sh = np.stack([rng.permutation(N) for _ in range(trials)]).astype(np.uint64)
    ro = oracle.Model(X, Y, method, A).cv_LSO(sh, test_size)
    rg = api.Model(X, Y, method, A).cv_LSO(sh, test_size)
    co = np.stack([e.T for e in ro.errors()])
    np.testing.assert_allclose(rg.cube, co, rtol=1e-10, atol=1e-10 * np.abs(co).max())
    np.testing.assert_allclose(rg.validation(api.RESS), ro.validation(oracle.RESS), rtol=RTOL)
    assert list(rg.optimal_num_components()) == [int(v) for v in ro.optimal_num_components()]


def test_selection_at_the_threshold_needs_exact_ranks(api, oracle):
    """alpha placed a hair below / above one test's p-value: the rank-sum bracket cannot decide, the exact sort must,
    and the decision has to flip exactly where the oracle's does (pls.cpp:283)."""
    par, met, _ = synth.make_set(6000, 4, 14, seed=123)
    X = oracle.colwise_z_scores(met); Y = oracle.colwise_z_scores(par)
    g = api.Model(X[:3000], Y[:3000]); o = oracle.Model(X[:3000], Y[:3000])
    res = o.cv_NEW_DATA(X[3000:], Y[3000:])
    E = res.errors()
    ref = np.argmin(res.validation(oracle.RESS), axis=1)
    ctx = api.get_context(0)
    checked = 0
    for y in range(4):
        if ref[y] == 0:
            continue
        ps = np.array([oracle.wilcoxon(E[y][:, ref[y]], E[y][:, alt]) for alt in range(ref[y])])
        for alt in {int(np.argmax(ps)), int(np.argmin(np.abs(ps - 0.1)))}:
            for alpha in (ps[alt] * (1 - 1e-12), ps[alt] * (1 + 1e-12)):
                if not (0 < alpha < 1):
                    continue
                n0 = ctx.exact_tests
                _, ncomp = g.cv_NEW_DATA(X[3000:], Y[3000:], alpha=alpha)
                assert list(ncomp) == [int(v) for v in res.optimal_num_components(alpha)], (y, alt, alpha)
                checked += ctx.exact_tests - n0
    assert checked > 0      # at least one of these decisions went through the exact path


def test_exact_level_from_fine_bins_matches_radix_and_oracle(api, oracle, monkeypatch):
    """Hold-out set large enough that a fine bin of level 2 holds dozens of elements (several 32-element chunks per bin, the crowded
    clamped tail bin) and a tenth of the rows duplicated (equal keys inside a bin): alpha a hair either side of a p-value forces the
    exact level; its two implementations (ranking inside the fine bins / radix sort of all keys) and the oracle must agree."""
    n_tr, n_te = 4000, 260003
    par, met, _ = synth.make_set(n_tr + n_te, 3, 7, seed=777)
    dup = np.arange(n_tr + 5, n_tr + n_te, 10)
    met[dup] = met[dup - 3]; par[dup] = par[dup - 3]
    X = oracle.colwise_z_scores(met); Y = oracle.colwise_z_scores(par)
    g = api.Model(X[:n_tr], Y[:n_tr]); o = oracle.Model(X[:n_tr], Y[:n_tr])
    res = o.cv_NEW_DATA(X[n_tr:], Y[n_tr:])
    E = res.errors()
    ref = np.argmin(res.validation(oracle.RESS), axis=1)
    ctx = api.get_context(0)
    exact = {False: 0, True: 0}
    for y in range(3):
        if ref[y] == 0:
            continue
        ps = np.array([oracle.wilcoxon(E[y][:, ref[y]], E[y][:, alt]) for alt in range(ref[y])])
        alt = int(np.argmin(np.abs(ps - 0.1)))
        for alpha in (ps[alt] * (1 - 1e-12), ps[alt] * (1 + 1e-12)):
            if not (0 < alpha < 1):
                continue
            want = [int(v) for v in res.optimal_num_components(alpha)]
            for radix in (False, True):
                if radix:
                    monkeypatch.setenv("ABCB200_EXACT_RADIX", "1")
                else:
                    monkeypatch.delenv("ABCB200_EXACT_RADIX", raising=False)
                n0, r0 = ctx.exact_tests, ctx.stat(7)
                _, ncomp = g.cv_NEW_DATA(X[n_tr:], Y[n_tr:], alpha=alpha)
                assert list(ncomp) == want, (y, alt, alpha, radix)
                exact[radix] += ctx.exact_tests - n0
                assert (ctx.stat(7) - r0 > 0) == (radix and ctx.exact_tests > n0), "the fine-bin level fell back to the radix sort"
    monkeypatch.delenv("ABCB200_EXACT_RADIX", raising=False)
    assert exact[False] > 0 and exact[False] == exact[True]


def test_selection_large_holdout_matches_oracle(api, oracle):
    """more responses and components than one level-1 CTA group; odd sizes"""
    par, met, _ = synth.make_set(30011, 7, 27, seed=321)
    X = oracle.colwise_z_scores(met); Y = oracle.colwise_z_scores(par)
    g = api.Model(X[:15000], Y[:15000]); o = oracle.Model(X[:15000], Y[:15000])
    res = o.cv_NEW_DATA(X[15000:], Y[15000:])
    for alpha in (0.1, 0.5, 0.01, 0.9):
        press, ncomp = g.cv_NEW_DATA(X[15000:], Y[15000:], alpha=alpha)
        np.testing.assert_allclose(press, res.validation(oracle.RESS), rtol=RTOL)
        assert list(ncomp) == [int(v) for v in res.optimal_num_components(alpha)], alpha


# ---- ranking ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,f", [((2000, 3, 6), 0.5), ((5001, 10, 20), 0.5), ((6000, 4, 8), 0.7), ((8000, 30, 40), 0.5), ((4100, 5, 33), 0.5), ((3999, 6, 37), 0.6), ((6100, 7, 72), 0.5)])
@pytest.mark.parametrize("method", [0, 1, 2])
def test_particle_ranking_pls(api, oracle, shape, f, method):
    N, P, K = shape
    par, met, target = synth.make_set(N, P, K, seed=1000 + N)
    o = oracle.particle_ranking_PLS(met, par, target, f)
    g = api.particle_ranking_PLS(met, par, target, f, method=method, return_info=True)
    assert g["ncomp_used"] == o["ncomp_used"]
    assert list(g["ncomp"]) == [int(v) for v in o["ncomp"]]
    np.testing.assert_allclose(g["dist"], o["dist"], rtol=RTOL)
    _assert_order_parity(g["order"], o["order"], o["dist"], N)
    top = api.particle_ranking_PLS(met, par, target, f, top_n=100, method=method)
    assert np.array_equal(top, g["order"][:100])


@pytest.mark.parametrize("shape", [(3000, 20, 180), (2400, 12, 300), (4200, 8, 500)])
def test_particle_ranking_pls_wide(api, oracle, shape):
    """wide metric blocks: K = 180 is the largest class of the all-on-chip component loop (pls_defl.cu), K = 300 and
    K = 500 (config 5's width) take the L2-streamed loop (pls_gram.cu)"""
    N, P, K = shape
    par, met, target = synth.make_set(N, P, K, seed=1000 + N)
    o = oracle.particle_ranking_PLS(met, par, target, 0.5)
    g = api.particle_ranking_PLS(met, par, target, 0.5, return_info=True)
    assert g["ncomp_used"] == o["ncomp_used"]
    assert list(g["ncomp"]) == [int(v) for v in o["ncomp"]]
    np.testing.assert_allclose(g["dist"], o["dist"], rtol=RTOL)
    _assert_order_parity(g["order"], o["order"], o["dist"], N)


def test_particle_ranking_simple(api, oracle):
    par, met, target = synth.make_set(30000, 3, 6, seed=41)
    o = oracle.particle_ranking_simple(met, target)
    g = api.particle_ranking_simple(met, par, target, return_info=True)
    np.testing.assert_allclose(g["dist"], o["dist"], rtol=RTOL)
    _assert_order_parity(g["order"], o["order"], o["dist"], 30000)


def test_ranking_rejects_bad_arguments(api):
    par, met, target = synth.make_set(100, 3, 6, seed=1)
    with pytest.raises(Exception):
        api.particle_ranking_PLS(met, par, target, 0.0)
    with pytest.raises(Exception):
        api.particle_ranking_PLS(met, par, target, 1.5)
    with pytest.raises(Exception):
        api.particle_ranking_PLS(met[:8], par[:8], target, 0.5)       # fewer training rows than components


# ---- weights ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(300, 200, 5), (1000, 1000, 10), (777, 1300, 30), (512, 640, 50), (100, 90, 70)])
@pytest.mark.parametrize("algo", [0, 1, 2])
def test_weights(api, oracle, shape, algo):
    n_new, n_old, P = shape
    if algo == 2 and P + 2 > 64:
        pytest.skip("DMMA kernel covers P <= 62")
    th_new, th_old, w_old, dv = synth.make_weight_case(n_new, n_old, P, seed=51 + P)
    rng = np.random.default_rng(P)
    numer = rng.uniform(0.5, 2.0, n_new)
    w = api.weight_predictive_prior(numer, th_new, th_old, w_old, dv, algo=algo)
    np.testing.assert_allclose(w, oracle.weight_predictive_prior(numer, th_new, th_old, w_old, dv), rtol=RTOL)
    assert abs(np.sum(w * w) - 1.0) < 1e-12
    np.testing.assert_allclose(api.weight_predictive_prior(None, th_new[:50]), np.full(50, 1 / 50))


def test_weights_set0_feeds_set1(api, oracle):
    th_new, th_old, _, dv = synth.make_weight_case(400, 300, 4, seed=9)
    w0 = api.weight_predictive_prior(None, th_old)                      # uniform 1/N, not normalised (AbcUtil.cpp:539-545)
    assert np.all(w0 == 1.0 / 300)
    w = api.weight_predictive_prior(None, th_new, th_old, w0, dv)
    np.testing.assert_allclose(w, oracle.weight_predictive_prior(np.ones(400), th_new, th_old, w0, dv), rtol=RTOL)


@pytest.mark.parametrize("algo", [1, 2])
def test_weights_converged_parameter(api, oracle, algo):
    th_new, th_old, w_old, dv = synth.make_weight_case(50, 40, 3, seed=61)
    th_new[:, 1] = 0.25; th_old[:, 1] = 0.25; dv[1] = 0.0               # dv == 0, equal values: factor skipped (:573)
    w = api.weight_predictive_prior(None, th_new, th_old, w_old, dv, algo=algo)
    np.testing.assert_allclose(w, oracle.weight_predictive_prior(np.ones(50), th_new, th_old, w_old, dv), rtol=RTOL)
    th_new[3, 1] = 0.5                                                  # differing value: NaN row, vector left unscaled
    w = api.weight_predictive_prior(None, th_new, th_old, w_old, dv, algo=algo)
    wo = oracle.weight_predictive_prior(np.ones(50), th_new, th_old, w_old, dv)
    assert np.isnan(w[3]) and np.isnan(wo[3])
    keep = np.arange(50) != 3
    np.testing.assert_allclose(w[keep], wo[keep], rtol=RTOL)


def test_weights_ill_conditioned_falls_back(api, oracle):
    # parameters far from the centre relative to the kernel bandwidth: the expanded form would lose digits,
    # algo=0 must pick the pairwise-difference kernel and stay within tolerance
    rng = np.random.default_rng(4)
    th_old = np.asfortranarray(1e4 + rng.standard_normal((200, 4)) * np.array([1.0, 50.0, 1e3, 1e-2]))
    th_new = np.asfortranarray(th_old[rng.integers(0, 200, 150)] + 0.01 * rng.standard_normal((150, 4)))
    dv = np.full(4, 2e-4); w_old = np.full(200, 1 / 200)
    w = api.weight_predictive_prior(None, th_new, th_old, w_old, dv, algo=0)
    np.testing.assert_allclose(w, oracle.weight_predictive_prior(np.ones(150), th_new, th_old, w_old, dv), rtol=1e-10)


def test_full_set_flow_like_abcsmc(api, oracle):
    """AbcSmc.cpp:634-664 + 1041-1066: rank -> truncate -> gather -> doubled variance -> weights."""
    cfg = synth.make_config("C2", scale=0.05)
    o = oracle.particle_ranking_PLS(cfg["metrics"], cfg["params"], cfg["target"], 0.5)
    order = api.particle_ranking_PLS(cfg["metrics"], cfg["params"], cfg["target"], 0.5, top_n=cfg["N_pp"])
    _assert_order_parity(order, o["order"], o["dist"], cfg["N_pp"])
    sel = cfg["params"][order.astype(np.int64), :]
    dv = api.calculate_doubled_variance(sel)
    np.testing.assert_allclose(dv, oracle.calculate_doubled_variance(sel), rtol=RTOL)
    w = api.weight_predictive_prior(None, sel, cfg["theta_old"], cfg["w_old"], cfg["dv_old"])
    np.testing.assert_allclose(w, oracle.weight_predictive_prior(np.ones(len(sel)), sel, cfg["theta_old"], cfg["w_old"], cfg["dv_old"]), rtol=RTOL)
