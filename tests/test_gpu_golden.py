"""Full-size parity against committed oracle outputs (tests/golden/fullsize_*.npz, made on CPU by
tests/golden/make_fullsize_goldens.py from oracle/abc_oracle.cpp: minutes to an hour per file, so the GPU box only reads them).
Inputs are regenerated from the same seeds (abcsmc_b200/synth.py). Bars (BASELINE.json north_star): selected indices, their
order and the component counts bit-exact; PRESS, distances, doubled variance and normalised weights within 1e-10 relative.
  C3   = configs[2], N=250k, K=150, P=30, top-N 5k (full size)
  T1M  = the north-star target shape, N=1M, K=150, P=30, top-N 10k (full size)
  C5   = configs[4] shape K=500, P=50 at N=200k (the oracle's residual cube at 1M is 100 GB)
  C4rows = 64 new-particle rows of the 1M x 1M x 30 weight update
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


def _load(tag):
    path = os.path.join(GOLD, f"fullsize_{tag}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated (tests/golden/make_fullsize_goldens.py)")
    return np.load(path)


@pytest.mark.parametrize("tag,name,scale", [("C3", "C3", 1.0), ("T1M", "T1M", 1.0), ("C5_s0.2", "C5", 0.2)])
def test_full_step_matches_oracle_golden(api, tag, name, scale):
    """The reference's call sequence (AbcSmc.cpp:634-664, 1041-1066) through the host C ABI against the oracle's outputs."""
    g = _load(tag)
    cfg = synth.make_config(name, scale=scale)
    N, N_pp = cfg["N"], cfg["N_pp"]
    assert (N, cfg["K"], cfg["P"], N_pp) == (int(g["N"]), int(g["K"]), int(g["P"]), int(g["N_pp"]))
    r = api.particle_ranking_PLS(cfg["metrics"], cfg["params"], cfg["target"], 0.5, top_n=N_pp, return_info=True)
    # discrete outputs: bit-exact
    assert list(r["ncomp"]) == [int(v) for v in g["ncomp"]]
    assert r["ncomp_used"] == int(g["ncomp_used"])
    order = r["order"].astype(np.int64)
    gold_order = g["order_top"].astype(np.int64)
    assert float(g["min_rel_gap_top"]) > 1e-11           # the oracle's top-N is not within rounding of a tie: exact equality is owed
    assert np.array_equal(order, gold_order)
    # distances: the selected ones, 4096 sampled ones, and the sum over all N
    np.testing.assert_allclose(r["dist"][gold_order], g["dist_top"], rtol=RTOL)
    np.testing.assert_allclose(r["dist"][g["dist_sample_idx"]], g["dist_sample"], rtol=RTOL)
    np.testing.assert_allclose(np.sum(r["dist"]), float(g["dist_sum"]), rtol=RTOL)
    # doubled variance of the selected rows in rank order, then the weight update against the previous predictive prior
    sel = np.asfortranarray(cfg["params"][order, :])
    np.testing.assert_allclose(api.calculate_doubled_variance(sel), g["dv"], rtol=RTOL)
    w = api.weight_predictive_prior(None, sel, cfg["theta_old"], cfg["w_old"], cfg["dv_old"])
    np.testing.assert_allclose(w, g["w"], rtol=RTOL)


@pytest.mark.parametrize("tag,name,scale", [("C3", "C3", 1.0)])
def test_press_matches_oracle_golden(api, tag, name, scale):
    """PLS::validation(RESS) (pls.cpp:235-261) of the hold-out half at the full C3 size, through the Model API."""
    g = _load(tag)
    cfg = synth.make_config(name, scale=scale)
    N = cfg["N"]
    X = api.colwise_z_scores(cfg["metrics"]); Y = api.colwise_z_scores(cfg["params"])
    n_tr = int(np.floor(N * 0.5 + 0.5))
    m = api.Model(X[:n_tr], Y[:n_tr])
    press, ncomp = m.cv_NEW_DATA(X[n_tr:], Y[n_tr:], alpha=0.1)
    np.testing.assert_allclose(press, g["press"], rtol=RTOL)
    assert list(ncomp) == [int(v) for v in g["ncomp"]]


@pytest.mark.parametrize("algo", [0, 1, 2])
def test_c4_rows_match_oracle_golden(api, algo):
    """64 rows of the C4 stress shape against all 1M old particles (src/AbcUtil.cpp:556-581): the expanded DMMA form with the
    centre at the first old particle, the pairwise-difference kernel, and the device-side choice between them."""
    g = _load("C4rows")
    c = synth.CONFIGS["C4"]
    th_new, th_old, w_old, dv_old = synth.make_weight_case(c["N_new"], c["N_old"], c["P"], c["seed"])
    sub = np.asfortranarray(th_new[g["rows"], :])
    w = api.weight_predictive_prior(None, sub, th_old, w_old, dv_old, algo=algo)
    np.testing.assert_allclose(w, g["w_normalised_over_rows"], rtol=RTOL)


def test_sharded_slices_match_full_and_oracle(api, oracle):
    """The multi-GPU decomposition on ONE device: two row slices (odd split) through abcb200_weights_unnorm_dev, the sum of
    squares added by hand (what the all-reduce does), abcb200_scale_weights_dev per slice — against abcb200_weights_dev on
    all rows and against the oracle. The slices are addressed as the sharded path does (pointer offset, ld = N_new)."""
    import ctypes as C
    import torch
    from abcsmc_b200 import device as dev
    ctx = api.get_context(0)
    dev.use_torch_stream(ctx)
    n_new, n_old, P = 3001, 2000, 30
    th_new, th_old, w_old, dv_old = synth.make_weight_case(n_new, n_old, P, 0xABC5F00D)
    numer = 0.5 + synth.uniform(0xABC5F00D, n_new, 77)
    d = torch.device("cuda", 0)
    t_new = dev.host_to_colmajor_tensor(th_new, d); t_old = dev.host_to_colmajor_tensor(th_old, d)
    t_w = torch.from_numpy(w_old).to(d); t_dv = torch.from_numpy(dv_old).to(d); t_num = torch.from_numpy(numer).to(d)
    want = oracle.weight_predictive_prior(numer, th_new, th_old, w_old, dv_old)
    for algo in (0, 1, 2):
        full = dev.weights(ctx, t_num, t_new, t_old, t_w, t_dv, algo=algo).cpu().numpy()
        np.testing.assert_allclose(full, want, rtol=RTOL)
        bounds = [(0, 1501), (1501, n_new)]
        parts, ss = [], []
        for lo, hi in bounds:
            w_loc = torch.zeros(hi - lo, dtype=torch.float64, device=d)
            s = torch.zeros(1, dtype=torch.float64, device=d)
            ctx.check(ctx._lib.abcb200_weights_unnorm_dev(ctx._h, C.c_void_p(t_num.data_ptr() + 8 * lo), C.c_void_p(t_new.data_ptr() + 8 * lo), n_new, hi - lo,
                                                          C.c_void_p(t_old.data_ptr()), n_old, n_old, C.c_void_p(t_w.data_ptr()), C.c_void_p(t_dv.data_ptr()),
                                                          P, algo, C.c_void_p(w_loc.data_ptr()), C.c_void_p(s.data_ptr())))
            parts.append(w_loc); ss.append(s)
        total = ss[0] + ss[1]                                   # the all-reduce
        for (lo, hi), w_loc in zip(bounds, parts):
            ctx.check(ctx._lib.abcb200_scale_weights_dev(ctx._h, C.c_void_p(w_loc.data_ptr()), hi - lo, C.c_void_p(total.data_ptr())))
        torch.cuda.synchronize()
        got = torch.cat(parts).cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=RTOL)
        np.testing.assert_allclose(got, full, rtol=1e-13)


# ---- against the REFERENCE'S OWN CODE (tests/golden/ref_*.npz, written by make_ref_fixtures.py from oracle/_ref/libabcref.so: the
# reference's unmodified pls.cpp + AbcUtil.cpp compiled against the Eigen / GSL stand-ins of oracle/shim/) -------------------------
def _ref_fixture(name):
    path = os.path.join(GOLD, name)
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated (tests/golden/make_ref_fixtures.py)")
    return np.load(path)


@pytest.mark.parametrize("tag,name", [("C2s", "C2"), ("C3s", "C3")])
def test_full_step_matches_reference_code_fixture(api, tag, name):
    """The CUDA path against ABC::particle_ranking_PLS / calculate_doubled_variance / weight_predictive_prior as the reference's own
    source computes them (no oracle in between): the whole order, component counts, PRESS, distances, dv, weights."""
    fx = _ref_fixture("ref_small.npz")
    N, K, P, n_pp = (int(v) for v in fx[f"{tag}_shape"])
    cfg = synth.make_config(name, scale=float(fx[f"{tag}_scale"]))
    assert (cfg["N"], cfg["K"], cfg["P"], cfg["N_pp"]) == (N, K, P, n_pp)
    r = api.particle_ranking_PLS(cfg["metrics"], cfg["params"], cfg["target"], 0.5, top_n=0, return_info=True)
    assert list(r["ncomp"]) == [int(v) for v in fx[f"{tag}_ncomp"]] and r["ncomp_used"] == int(fx[f"{tag}_ncomp_used"])
    ref_dist = fx[f"{tag}_dist"]
    np.testing.assert_allclose(r["dist"], ref_dist, rtol=RTOL)
    order = r["order"].astype(np.int64); ref_order = fx[f"{tag}_order"].astype(np.int64)
    assert np.array_equal(order, ref_order)
    top = order[:n_pp]
    sel = np.asfortranarray(cfg["params"][top, :])
    np.testing.assert_allclose(api.calculate_doubled_variance(sel), fx[f"{tag}_dv"], rtol=RTOL)
    np.testing.assert_allclose(api.weight_predictive_prior(None, sel, cfg["theta_old"], cfg["w_old"], cfg["dv_old"]), fx[f"{tag}_w"], rtol=RTOL)
    X = api.colwise_z_scores(cfg["metrics"]); Y = api.colwise_z_scores(cfg["params"])
    n_tr = int(np.floor(N * 0.5 + 0.5))
    m = api.Model(X[:n_tr], Y[:n_tr])
    press, ncomp = m.cv_NEW_DATA(X[n_tr:], Y[n_tr:], alpha=0.1)
    np.testing.assert_allclose(press, fx[f"{tag}_press"], rtol=RTOL)
    assert list(ncomp) == [int(v) for v in fx[f"{tag}_ncomp"]]
    Rg, Rr = m.R, fx[f"{tag}_R"]
    s = np.sign(np.sum(Rg * Rr, axis=0))
    np.testing.assert_allclose(Rg * s, Rr, rtol=0, atol=RTOL * np.abs(Rr).max())
    used = int(fx[f"{tag}_ncomp_used"])
    np.testing.assert_allclose(m.coefficients(used), fx[f"{tag}_coef"], rtol=0, atol=RTOL * np.abs(fx[f"{tag}_coef"]).max())


@pytest.mark.parametrize("tag", ["toy", "nir"])
@pytest.mark.parametrize("method", [0, 1])
def test_model_matches_reference_code_fixture(api, tag, method):
    """PLS::Model on the reference's demo inputs (lib/PLS/src/main.cpp:19-41) against the reference's own code: coefficients, cv_LOO,
    cv_LSO on the reference's mt19937 partitions; nir / octane is the single-response branch (pls.cpp:403-404)."""
    fx = _ref_fixture("ref_small.npz")
    d = np.load(os.path.join(GOLD, "toy_inputs.npz"))
    X = api.colwise_z_scores(d["toyX"] if tag == "toy" else d["nir"]); Y = api.colwise_z_scores(d["toyY"] if tag == "toy" else d["octane"].reshape(-1, 1))
    A = int(fx[f"{tag}_A"]); key = f"{tag}_m{method}"
    m = api.Model(X, Y, method, A)
    want = fx[f"{key}_coef"]
    np.testing.assert_allclose(m.coefficients(), want, rtol=0, atol=RTOL * np.abs(want).max())
    np.testing.assert_allclose(m.explained_variance(X, Y), fx[f"{key}_ev"], rtol=0, atol=RTOL)
    loo = m.cv_LOO()
    np.testing.assert_allclose(loo.validation(api.RESS), fx[f"{key}_loo_press"], rtol=1e-9)
    assert list(loo.optimal_num_components(0.1)) == [int(v) for v in fx[f"{key}_loo_ncomp"]]
    n = X.shape[0]
    lso = m.cv_LSO(fx[f"{key}_lso_shuffles"], int(0.3 * n + 0.5))
    np.testing.assert_allclose(lso.validation(api.RESS), fx[f"{key}_lso_press"], rtol=1e-9)


@pytest.mark.parametrize("name", ["C3", "C2"])
def test_fullsize_order_matches_reference_code(api, name):
    """Full dengue shape (C3: N=250k, K=150, P=30, first 5000 ranks) and configs[1] (C2: N=100k, K=20, P=10, first 1000 ranks) against
    the order ABC::particle_ranking_PLS returned when the reference's own source was run on the same inputs
    (tests/golden/ref_fullsize_<name>.npz)."""
    g = _ref_fixture(f"ref_fullsize_{name}.npz")
    cfg = synth.make_config(name, scale=1.0)
    assert (cfg["N"], cfg["K"], cfg["P"], cfg["N_pp"]) == (int(g["N"]), int(g["K"]), int(g["P"]), int(g["N_pp"]))
    r = api.particle_ranking_PLS(cfg["metrics"], cfg["params"], cfg["target"], 0.5, top_n=cfg["N_pp"], return_info=True)
    assert np.array_equal(r["order"].astype(np.int64), g["order_top"].astype(np.int64))


@pytest.mark.parametrize("tag", ["dengue_sqlite", "dengue_pp250"])
def test_full_step_on_the_reference_s_own_data(api, tag):
    """The CUDA path on the particle sets the reference ships (examples/scratch/posterior.sqlite: 1000 x 5 x 7 of a dengue-model fit;
    vis/dengue_predictive_prior-full_ts.06: 250 x 4 x 6; values at AbcSmc's 6 significant digits) against the outputs of the reference's
    own code on them (tests/golden/ref_realdata.npz): order, component counts, distances, PRESS, doubled variance, weights.
    The oracle is held to the same fixture on the CPU (tests/test_ref_pin.py). tools/realdata_check.py is the same comparison without pytest:
    on the B200 both sets gave identical orders and component counts, distances 3e-15 / 2e-14, doubled variance 5e-16, weights 1e-15."""
    g = _ref_fixture("ref_realdata.npz")
    met, par, target = np.asfortranarray(g[f"{tag}_met"]), np.asfortranarray(g[f"{tag}_par"]), g[f"{tag}_target"]
    r = api.particle_ranking_PLS(met, par, target, 0.5, top_n=0, return_info=True)
    assert list(r["ncomp"]) == [int(v) for v in g[f"{tag}_ncomp"]] and r["ncomp_used"] == int(g[f"{tag}_ncomp_used"])
    np.testing.assert_allclose(r["dist"], g[f"{tag}_dist"], rtol=RTOL)
    order = r["order"].astype(np.int64)
    assert np.array_equal(order, g[f"{tag}_order"].astype(np.int64))
    n_pp = met.shape[0] // 10
    th_new, th_old = np.asfortranarray(par[order[:n_pp]]), np.asfortranarray(par[order[n_pp:2 * n_pp]])
    np.testing.assert_allclose(api.calculate_doubled_variance(th_new), g[f"{tag}_dv"], rtol=RTOL)
    dv_old = api.calculate_doubled_variance(th_old)
    np.testing.assert_allclose(dv_old, g[f"{tag}_dv_next"], rtol=RTOL)
    numer = np.full(n_pp, np.prod(1.0 / (g[f"{tag}_prior_hi"] - g[f"{tag}_prior_lo"])))
    w = api.weight_predictive_prior(numer, th_new, th_old, np.full(n_pp, 1.0 / n_pp), g[f"{tag}_dv_next"])
    np.testing.assert_allclose(w, g[f"{tag}_w_vs_next"], rtol=RTOL)
