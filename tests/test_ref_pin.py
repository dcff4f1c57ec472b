"""Pins oracle/abc_oracle.cpp (the hand restatement every GPU parity test uses as its checker) to the REFERENCE'S OWN CODE.

oracle/_ref/libabcref.so = the reference's unmodified lib/PLS/src/pls.cpp + src/AbcUtil.cpp compiled where they lie under
/root/reference against the Eigen / GSL stand-ins of oracle/shim/ (oracle/Makefile `ref`, oracle/ref_harness.cpp): the reference's
statements run, Eigen's kernels are naive loops, the eigen-solver is tred2 / tql2 (the oracle uses cyclic Jacobi).

Two layers:
  * fixtures (run everywhere, the GPU box included): tests/golden/ref_small.npz and ref_fullsize_C3.npz were written by
    tests/golden/make_ref_fixtures.py from that library; the oracle must reproduce them — discrete outputs (orders, component
    counts) bit-exactly, FP64 outputs to 1e-10 (the bar north_star sets; observed <= 1e-13);
  * live (only where /root/reference exists, i.e. the authoring container): the same comparison function by function on fresh
    seeded inputs, cv_LOO / cv_LSO included.
CPU only: nothing here touches the GPU or the product library."""
import os

import numpy as np
import pytest

from abcsmc_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-10


def _align(a, like):
    """PLS factors are defined up to the sign of each component's eigenvector (Eigen's, the stand-in's and the oracle's differ)."""
    s = np.sign(np.sum(a * like, axis=0)); s[s == 0] = 1.0
    return a * s


def _close(got, want, rtol=RTOL):
    np.testing.assert_allclose(got, want, rtol=0, atol=rtol * max(np.abs(want).max(), 1e-300))


@pytest.fixture(scope="module")
def fx():
    path = os.path.join(GOLD, "ref_small.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/ref_small.npz not generated (tests/golden/make_ref_fixtures.py small)")
    return np.load(path)


# ---- layer 1: the oracle against the committed outputs of the reference's own code ---------------------------------------------
@pytest.mark.parametrize("tag,name", [("C2s", "C2"), ("C3s", "C3")])
def test_oracle_path_matches_reference_fixture(oracle, fx, tag, name):
    """The per-set sequence (AbcUtil.cpp:423-458, AbcSmc.cpp:645-664, 1041-1066) at the two AbcSmc shapes, scaled down."""
    N, K, P, n_pp = (int(v) for v in fx[f"{tag}_shape"])
    cfg = synth.make_config(name, scale=float(fx[f"{tag}_scale"]))
    assert (cfg["N"], cfg["K"], cfg["P"], cfg["N_pp"]) == (N, K, P, n_pp)
    r = oracle.particle_ranking_PLS(cfg["metrics"], cfg["params"], cfg["target"], 0.5)
    assert np.array_equal(r["order"], fx[f"{tag}_order"])                       # the whole order, bit-exact
    assert np.array_equal(r["ncomp"], fx[f"{tag}_ncomp"]) and r["ncomp_used"] == int(fx[f"{tag}_ncomp_used"])
    _close(r["press"], fx[f"{tag}_press"])
    np.testing.assert_allclose(r["dist"], fx[f"{tag}_dist"], rtol=RTOL)
    mean = oracle.colwise_mean(cfg["metrics"])
    assert np.array_equal(mean, fx[f"{tag}_mean"]) and np.array_equal(oracle.colwise_stdev(cfg["metrics"], mean), fx[f"{tag}_sd"])
    top = r["order"][:n_pp].astype(np.int64)
    sel = np.asfortranarray(cfg["params"][top, :])
    np.testing.assert_allclose(oracle.calculate_doubled_variance(sel), fx[f"{tag}_dv"], rtol=1e-15)
    w = oracle.weight_predictive_prior(np.ones(n_pp), sel, cfg["theta_old"], cfg["w_old"], cfg["dv_old"])
    np.testing.assert_allclose(w, fx[f"{tag}_w"], rtol=RTOL)


@pytest.mark.parametrize("tag", ["toy", "nir"])
@pytest.mark.parametrize("method", [0, 1])
def test_oracle_model_matches_reference_fixture(oracle, fx, tag, method):
    """PLS::Model on the reference's demo inputs (lib/PLS/src/main.cpp:19-41): factors, coefficients, SSE, explained variance, cv_LOO,
    cv_LSO with the reference's own mt19937 partitions; nir / octane is the M == 1 branch (pls.cpp:403-404)."""
    d = np.load(os.path.join(GOLD, "toy_inputs.npz"))
    X = oracle.colwise_z_scores(d["toyX"] if tag == "toy" else d["nir"]); Y = oracle.colwise_z_scores(d["toyY"] if tag == "toy" else d["octane"])
    A = int(fx[f"{tag}_A"]); key = f"{tag}_m{method}"
    m = oracle.Model(X, Y, method, A)
    for name in ("R", "P", "W", "Q"):
        want = fx[f"{key}_{name}"]
        _close(_align(getattr(m, name), want), want)
    _close(m.coefficients(), fx[f"{key}_coef"])
    np.testing.assert_allclose(m.SSE(X, Y), fx[f"{key}_sse"], rtol=1e-9)       # SSE of a near-perfect fit: cancellation, not a path output
    _close(m.explained_variance(X, Y), fx[f"{key}_ev"])
    loo = m.cv_LOO()
    _close(loo.validation(oracle.RESS), fx[f"{key}_loo_press"])
    assert np.array_equal(loo.optimal_num_components(0.1), fx[f"{key}_loo_ncomp"])
    n = X.shape[0]; test_size = int(0.3 * n + 0.5)
    lso = m.cv_LSO(fx[f"{key}_lso_shuffles"], test_size)
    _close(lso.validation(oracle.RESS), fx[f"{key}_lso_press"])


def test_oracle_scalars_match_reference_fixture(oracle, fx):
    """wilcoxon / normalcdf (pls.cpp:152-211), prior likelihoods (Priors.h), the converged-parameter rule (AbcUtil.cpp:573), set-0
    weights (:539-545) — bit-exact: same statements, same libm."""
    e1, e2 = fx["wilcoxon_e1"], fx["wilcoxon_e2"]
    assert np.array_equal(np.array([oracle.wilcoxon(e1, e2), oracle.wilcoxon(e2, e1), oracle.wilcoxon(e1, e1)]), fx["wilcoxon_p"])
    assert np.array_equal(np.array([oracle.normalcdf(z) for z in fx["normalcdf_z"]]), fx["normalcdf"])
    vals = fx["prior_vals"]
    lik = np.array([[oracle.prior_likelihood(t, a, b, v) for v in vals] for t, a, b in ((0, 0.0, 2.0), (1, 0.0, 3.0), (2, 1.0, 0.5))])
    assert np.array_equal(lik, fx["prior_lik"])
    th_new, th_old = fx["wconv_th_new"], fx["wconv_th_old"]
    numer = np.array([oracle.prior_likelihood(0, 0.0, 1.0, r[0]) * oracle.prior_likelihood(0, 0.0, 1.0, r[1]) * oracle.prior_likelihood(2, 0.5, 0.3, r[2]) for r in th_new])
    assert fx["wconv_dv"][1] == 0.0
    np.testing.assert_allclose(oracle.weight_predictive_prior(numer, th_new, th_old, fx["wconv_w_old"], fx["wconv_dv"]), fx["wconv_w"], rtol=1e-14)
    assert np.array_equal(oracle.weight_predictive_prior0(7), fx["w0"])


@pytest.mark.parametrize("tag", ["dengue_sqlite", "dengue_pp250"])
def test_oracle_matches_reference_on_the_reference_s_own_data(oracle, tag):
    """Inputs: the particle sets the reference ships (examples/scratch/posterior.sqlite: 1000 particles x 5 parameters x 7 metrics of a
    dengue-model fit; vis/dengue_predictive_prior-full_ts.06: 250 x 4 x 6), values at the 6 significant digits AbcSmc stores. Outputs of
    the reference's own code on them (tests/golden/ref_realdata.npz, make_ref_fixtures.py realdata): the oracle reproduces the whole
    order and the component counts exactly, the FP64 outputs to 1e-10."""
    path = os.path.join(GOLD, "ref_realdata.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/ref_realdata.npz not generated (tests/golden/make_ref_fixtures.py realdata)")
    g = np.load(path)
    met, par, target = g[f"{tag}_met"], g[f"{tag}_par"], g[f"{tag}_target"]
    r = oracle.particle_ranking_PLS(met, par, target, 0.5)
    assert np.array_equal(r["order"], g[f"{tag}_order"])
    assert np.array_equal(np.asarray(r["ncomp"]), g[f"{tag}_ncomp"]) and r["ncomp_used"] == int(g[f"{tag}_ncomp_used"])
    _close(r["dist"], g[f"{tag}_dist"]); _close(r["press"], g[f"{tag}_press"])
    N = met.shape[0]; n_pp = N // 10
    order = r["order"].astype(np.int64)
    th_new, th_old = par[order[:n_pp]], par[order[n_pp:2 * n_pp]]
    assert np.array_equal(oracle.calculate_doubled_variance(th_new), g[f"{tag}_dv"])
    dv_old = oracle.calculate_doubled_variance(th_old)
    assert np.array_equal(dv_old, g[f"{tag}_dv_next"])
    numer = np.full(n_pp, np.prod(1.0 / (g[f"{tag}_prior_hi"] - g[f"{tag}_prior_lo"])))
    np.testing.assert_allclose(oracle.weight_predictive_prior(numer, th_new, th_old, np.full(n_pp, 1.0 / n_pp), dv_old), g[f"{tag}_w_vs_next"], rtol=1e-12)


def test_report_statistics_match_reference_fixture(fx):
    """The numpy statements tests/test_gpu_chain.py checks the on-device filtering-report statistics with, against
    ABC::calculate_nrmse (AbcUtil.cpp:326-345) and ABC::median (:46-61)."""
    from test_gpu_chain import _nrmse
    np.testing.assert_allclose(_nrmse(fx["nrmse_mets"], fx["nrmse_obs"]), float(fx["nrmse"]), rtol=1e-14)
    assert np.median(fx["median_in"]) == fx["median"][0] and np.median(fx["median_in"][:32]) == fx["median"][1]


def test_fullsize_c3_order_from_reference_equals_oracle_golden():
    """The first N_pp entries of ABC::particle_ranking_PLS's order at the FULL dengue shape (N=250k, K=150, P=30), computed by the
    reference's own code (ref_fullsize_C3.npz), are the ones the oracle wrote into fullsize_C3.npz — the file the GPU test
    test_gpu_golden.py::test_full_step_matches_oracle_golden holds the CUDA path to, bit for bit."""
    a = os.path.join(GOLD, "ref_fullsize_C3.npz"); b = os.path.join(GOLD, "fullsize_C3.npz")
    if not (os.path.exists(a) and os.path.exists(b)):
        pytest.skip("full-size fixtures not generated")
    ra, rb = np.load(a), np.load(b)
    assert (int(ra["N"]), int(ra["K"]), int(ra["P"]), int(ra["N_pp"])) == (int(rb["N"]), int(rb["K"]), int(rb["P"]), int(rb["N_pp"]))
    assert np.array_equal(ra["order_top"].astype(np.int64), rb["order_top"].astype(np.int64))


def test_fullsize_c2_order_from_reference_equals_oracle(oracle):
    """configs[1] at full size (N=100k, K=20, P=10): the oracle's top-N (run here, ~2 s) against the reference's own code's."""
    a = os.path.join(GOLD, "ref_fullsize_C2.npz")
    if not os.path.exists(a):
        pytest.skip("full-size fixture not generated")
    g = np.load(a)
    cfg = synth.make_config("C2", scale=1.0)
    assert (cfg["N"], cfg["K"], cfg["P"], cfg["N_pp"]) == (int(g["N"]), int(g["K"]), int(g["P"]), int(g["N_pp"]))
    r = oracle.particle_ranking_PLS(cfg["metrics"], cfg["params"], cfg["target"], 0.5)
    assert np.array_equal(r["order"][:cfg["N_pp"]], g["order_top"])
    assert np.uint64(np.bitwise_xor.reduce(r["order"] * np.arange(1, r["order"].size + 1, dtype=np.uint64))) == g["order_checksum"]   # all 100k ranks


# ---- layer 2: live, where the reference's sources are present -------------------------------------------------------------------
@pytest.fixture(scope="module")
def ref():
    import oracle.ref as r
    if not os.path.exists(os.path.join(r.REFERENCE_ROOT, "lib", "PLS", "src", "pls.cpp")):
        pytest.skip("reference sources not present (GPU box): the committed fixtures above stand in")
    r.build()
    return r


@pytest.mark.parametrize("N,K,P,seed", [(900, 7, 3, 5), (1500, 33, 6, 6), (701, 12, 1, 7)])
def test_live_path_functions(oracle, ref, N, K, P, seed):
    par, met, target = synth.make_set(N, P, K, seed=seed)
    mean = ref.colwise_mean(met)
    assert np.array_equal(mean, oracle.colwise_mean(met))
    assert np.array_equal(ref.colwise_stdev(met, mean), oracle.colwise_stdev(met, mean))
    assert np.array_equal(ref.colwise_z_scores(met), oracle.colwise_z_scores(met))
    assert np.array_equal(ref.z_scores(target, mean, oracle.colwise_stdev(met, mean)), oracle.z_scores(target, mean, oracle.colwise_stdev(met, mean)))
    X = oracle.colwise_z_scores(met); Y = oracle.colwise_z_scores(par)
    n_tr = N // 2
    for method in (0, 1):
        mr = ref.Model(X[:n_tr], Y[:n_tr], method); mo = oracle.Model(X[:n_tr], Y[:n_tr], method)
        for name in ("R", "P", "W", "Q") + (("T",) if method == 0 else ()):
            want = getattr(mr, name)
            _close(_align(getattr(mo, name), want), want)
        for c in (1, K // 2 + 1, K):
            _close(mo.coefficients(c), mr.coefficients(c))
            _close(mo.fitted_values(X[n_tr:], c), mr.fitted_values(X[n_tr:], c))
        er = mr.cv_NEW_DATA(X[n_tr:], Y[n_tr:]); eo = mo.cv_NEW_DATA(X[n_tr:], Y[n_tr:])
        for a, b in zip(eo.errors(), er.errors()):
            _close(a, b)
        _close(eo.validation(oracle.RESS), er.validation(ref.RESS)); _close(eo.validation(oracle.MSE), er.validation(ref.MSE))
        for alpha in (0.1, 0.5, 0.01):
            assert np.array_equal(eo.optimal_num_components(alpha), er.optimal_num_components(alpha))
    ro = oracle.particle_ranking_PLS(met, par, target, 0.5)
    assert np.array_equal(ro["order"], ref.particle_ranking_PLS(met, par, target, 0.5))
    for frac in (0.3, 0.75):
        assert np.array_equal(oracle.particle_ranking_PLS(met, par, target, frac)["order"], ref.particle_ranking_PLS(met, par, target, frac))
    assert np.array_equal(oracle.particle_ranking_simple(met, target)["order"], ref.particle_ranking_simple(met, target))
    d = ro["dist"]
    assert np.array_equal(oracle.ordered(d), ref.ordered(d))
    S = X[:, :5]
    assert np.array_equal(oracle.euclidean(S, S[3]), ref.euclidean(S, S[3]))


@pytest.mark.parametrize("N,K,M", [(40, 6, 3), (25, 4, 1)])
def test_live_loo_lso(oracle, ref, N, K, M):
    par, met, _ = synth.make_set(N, M, K, seed=21)
    X = oracle.colwise_z_scores(met); Y = oracle.colwise_z_scores(par)
    for method in (0, 1):
        mr = ref.Model(X, Y, method, K - 1); mo = oracle.Model(X, Y, method, K - 1)
        lr, lo = mr.cv_LOO(), mo.cv_LOO()
        for a, b in zip(lo.errors(), lr.errors()):
            _close(a, b, 1e-9)                   # N - 1 = 24..39 training rows: the last components are ill-conditioned
        assert np.array_equal(lo.optimal_num_components(0.1), lr.optimal_num_components(0.1))
        test_size = int(0.25 * N + 0.5)
        sr = ref.cv_LSO_seeded(mr, 0.25, 6, 99)
        so = mo.cv_LSO(ref.lso_shuffles(99, N, test_size, 6), test_size)
        for a, b in zip(so.errors(), sr.errors()):
            _close(a, b, 1e-9)


def test_live_weights_and_variance(oracle, ref):
    rng = np.random.default_rng(3)
    P = 5
    th_old = rng.uniform(size=(211, P)); th_new = np.clip(th_old[rng.integers(0, 211, 173)] + 0.05 * rng.normal(size=(173, P)), -0.2, 1.2)
    th_old[:, 2] = np.round(th_old[:, 2] * 9); th_new[:, 2] = np.round(th_new[:, 2] * 9)         # an integer parameter
    w_old = rng.uniform(size=211); w_old /= np.linalg.norm(w_old)
    dv = ref.calculate_doubled_variance(th_old)
    assert np.array_equal(dv, oracle.calculate_doubled_variance(th_old))
    kinds = [(0, 0.0, 1.0), (2, 0.5, 0.4), (1, 0.0, 9.0), (0, -1.0, 2.0), (2, 0.0, 1.0)]
    pt, pa, pb = zip(*kinds)
    numer = np.ones(173)
    for p in range(P):
        numer *= np.array([oracle.prior_likelihood(pt[p], pa[p], pb[p], v) for v in th_new[:, p]])
    assert np.any(numer == 0.0)                                                                  # rows outside a uniform prior: weight 0
    wr = ref.weight_predictive_prior(pt, pa, pb, th_new, th_old, w_old, dv)
    wo = oracle.weight_predictive_prior(numer, th_new, th_old, w_old, dv)
    np.testing.assert_allclose(wo, wr, rtol=1e-14)
    assert np.array_equal(ref.weight_predictive_prior0(9, P), oracle.weight_predictive_prior0(9))
    e1 = rng.normal(size=2500); e2 = e1 * 1.003 + 0.01 * rng.normal(size=2500)
    assert ref.wilcoxon(e1, e2) == oracle.wilcoxon(e1, e2) and 0.0 < ref.wilcoxon(e1, e2) < 1.0
    for z in (-3.3, -0.4, 0.0, 0.7, 2.9):
        assert ref.normalcdf(z) == oracle.normalcdf(z)


def test_live_edge_semantics(oracle, ref):
    """SURVEY.md appendix A, statement by statement, on the reference's own code: the NaN quirk of colwise_z_scores (pls.cpp:94-103), a
    single row (colwise_stdev of N < 2, pls.cpp:69-87), an empty hold-out set (training_fraction = 1: PRESS all zero -> one component),
    duplicated rows (exact ties through Wilcoxon and the final ordering), the dv == 0 rule and the inf * 0 = NaN edge of the weight
    update (AbcUtil.cpp:573) with Eigen's normalize() leaving a vector that holds a NaN un-normalised, doubled variance of fewer than two rows (RunningStat.h)."""
    rng = np.random.default_rng(11)
    X = rng.normal(size=(50, 4)); X[:, 2] = 3.25                                   # a constant column
    zr, zo = ref.colwise_z_scores(X), oracle.colwise_z_scores(X)
    assert np.array_equal(np.isnan(zr), np.isnan(zo)) and np.all(np.isnan(zr[:, 2])) and np.array_equal(zr[:, [0, 1, 3]], zo[:, [0, 1, 3]])
    one = X[:1]
    sd1 = ref.colwise_stdev(one, ref.colwise_mean(one))                               # SST = 0 for N < 2 (:71), then sqrt(0 / 0) (:82)
    assert np.all(np.isnan(sd1)) and np.all(np.isnan(oracle.colwise_stdev(one, oracle.colwise_mean(one))))
    assert np.array_equal(ref.calculate_doubled_variance(X[:1]), oracle.calculate_doubled_variance(X[:1]))
    assert np.array_equal(ref.calculate_doubled_variance(X[:2]), oracle.calculate_doubled_variance(X[:2]))
    # training_fraction = 1: no hold-out rows
    par, met, target = synth.make_set(400, 3, 6, seed=31)
    assert np.array_equal(oracle.particle_ranking_PLS(met, par, target, 1.0)["order"], ref.particle_ranking_PLS(met, par, target, 1.0))
    assert oracle.particle_ranking_PLS(met, par, target, 1.0)["ncomp_used"] == 1
    # duplicated particles: ties inside the signed-rank tests and in the final order
    met2, par2 = met.copy(), par.copy()
    src = rng.integers(0, 400, 60); dst = rng.permutation(400)[:60]
    met2[dst] = met2[src]; par2[dst] = par2[src]
    ro = oracle.particle_ranking_PLS(met2, par2, target, 0.5)
    assert np.any(np.diff(np.sort(ro["dist"])) == 0)
    assert np.array_equal(ro["order"], ref.particle_ranking_PLS(met2, par2, target, 0.5))
    assert np.array_equal(oracle.particle_ranking_simple(met2, target)["order"], ref.particle_ranking_simple(met2, target))
    # weight update: a converged parameter (dv == 0) with equal values is skipped, with different values gives inf * 0 = NaN
    th_old = rng.uniform(size=(30, 3)); th_old[:, 1] = 0.5
    th_new = th_old[rng.integers(0, 30, 20)] + 0.01 * rng.normal(size=(20, 3)); th_new[:, 1] = 0.5
    w_old = rng.uniform(size=30); w_old /= np.linalg.norm(w_old)
    dv = oracle.calculate_doubled_variance(th_old)
    assert dv[1] == 0.0
    pt, pa, pb = [0, 0, 0], [-1.0, -1.0, -1.0], [2.0, 2.0, 2.0]
    numer = np.full(20, (1.0 / 3.0) ** 3)
    wr = ref.weight_predictive_prior(pt, pa, pb, th_new, th_old, w_old, dv)
    np.testing.assert_allclose(oracle.weight_predictive_prior(numer, th_new, th_old, w_old, dv), wr, rtol=1e-14)
    assert np.all(np.isfinite(wr))
    th_new[3, 1] = 0.75                                                              # differs from every old value while dv == 0
    wr = ref.weight_predictive_prior(pt, pa, pb, th_new, th_old, w_old, dv)
    wo = oracle.weight_predictive_prior(numer, th_new, th_old, w_old, dv)
    # squaredNorm() is NaN, Eigen's normalize() tests `z > 0` (false): row 3 is NaN and the rest is left UN-normalised (AbcUtil.cpp:583)
    assert np.array_equal(np.isnan(wr), np.isnan(wo)) and np.isnan(wr).sum() == 1 and np.isnan(wr[3])
    np.testing.assert_allclose(wo[~np.isnan(wo)], wr[~np.isnan(wr)], rtol=1e-14)
    assert np.linalg.norm(wr[~np.isnan(wr)]) < 0.5


def test_live_proposal_sampling(oracle, ref):
    """SURVEY.md §8 row f1 (AbcUtil.cpp:378-404, 462-488): the factor exactly; the reference's sampling statements, run on the
    stand-in's MT19937 stream, against the oracle's restatement on its own stream (distributional)."""
    from scipy import stats
    from test_sampling import _case, _mvn_case
    m = _mvn_case(n_pp=350, seed=21)
    L = ref.setup_mvn_sampler(m["theta"])
    np.testing.assert_allclose(oracle.setup_mvn_sampler(m["theta"]), L, rtol=1e-13, atol=1e-16)
    r = ref.sample_mvn_predictive_priors(3, 15000, m["w"], m["theta"], m["ptype"], m["pa"], m["pb"], L)
    o = oracle.sample_mvn_predictive_priors(4, 15000, m["w"], m["theta"], m["ptype"], m["pa"], m["pb"], L)["samples"]
    c = _case(n_pp=300, seed=22)
    r2 = ref.sample_predictive_priors(5, 15000, c["w"], c["theta"], c["ptype"], c["pa"], c["pb"], c["dv"])
    o2 = oracle.sample_predictive_priors(6, 15000, c["w"], c["theta"], c["ptype"], c["pa"], c["pb"], c["dv"])["samples"]
    for a, b in ((r, o), (r2, o2)):
        for p in range(3):
            assert stats.ks_2samp(a[:, p], b[:, p]).pvalue > 1e-4, p
    assert r[:, :2].min() >= 0.0 and r[:, :2].max() <= 1.0 and np.all(r2[:, 1] == np.round(r2[:, 1]))


def test_reference_own_test_program_passes_on_the_standins(ref):
    """/root/reference/tests/abcutil.cpp, unmodified, built on the reference's own pls.cpp + AbcUtil.cpp with the Eigen / GSL
    stand-ins (tests/cpp/Makefile): the reference's known answers hold for the library this file pins the oracle with."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-s", "-C", os.path.join(root, "tests", "cpp"), "_build/ref_tests_abcutil", "REF=" + ref.REFERENCE_ROOT])
    out = subprocess.run([os.path.join(root, "tests", "cpp", "_build", "ref_tests_abcutil")], capture_output=True, text=True, check=True).stdout
    assert out.count(" passed on line ") == 2 and "failed" not in out, out
