#!/usr/bin/env python
"""Fixtures from the REFERENCE'S OWN CODE (not from the oracle restatement).

oracle/_ref/libabcref.so is the reference's unmodified lib/PLS/src/pls.cpp and src/AbcUtil.cpp compiled where they lie under
/root/reference against the Eigen / GSL stand-ins of oracle/shim/ (oracle/Makefile target `ref`; oracle/ref_harness.cpp). It exists
only in the authoring container; this script stores its outputs so that they travel:

    python tests/golden/make_ref_fixtures.py small        # tests/golden/ref_small.npz     (~1 min)
    python tests/golden/make_ref_fixtures.py C3           # tests/golden/ref_fullsize_C3.npz (N=250k, K=150, P=30: ~6 min, ~20 GB)
    python tests/golden/make_ref_fixtures.py C2           # tests/golden/ref_fullsize_C2.npz (N=100k, K=20, P=10: seconds)
    python tests/golden/make_ref_fixtures.py realdata     # tests/golden/ref_realdata.npz: the reference's own data files as inputs
    python tests/golden/make_ref_fixtures.py sampling     # tests/golden/ref_sampling.npz: next-set proposals
    python tests/golden/make_ref_fixtures.py main         # tests/golden/ref_main_toy.txt: what the reference's demo PROGRAM prints
                                                          # (lib/PLS/src/main.cpp + pls.cpp, tests/cpp/Makefile) on toyX / toyY, 5 components

`small`: for each case the inputs come from abcsmc_b200/synth.py (seeded) or tests/golden/toy_inputs.npz (the reference's demo
files), the outputs from the reference's functions in the order AbcSmc calls them (src/AbcUtil.cpp:423-458, src/AbcSmc.cpp:634-664,
1041-1066): the full order from ABC::particle_ranking_PLS itself; and, through the same public calls that function makes
(colwise_z_scores -> Model -> cv_NEW_DATA -> validation / optimal_num_components -> scores -> euclidean), the quantities it does not
return: PRESS, component counts, R, distances; then calculate_doubled_variance and weight_predictive_prior on the top-N rows.
Full sizes: ABC::particle_ranking_PLS's order only (first N_pp entries kept).
`realdata`: the two particle sets the reference ships — examples/scratch/posterior.sqlite (1000 posterior particles of a dengue-model fit:
5 parameters, 7 metrics, old table names jobs / parameters / metrics) and vis/dengue_predictive_prior-full_ts.06 (250 ranked particles: 4
parameters, 6 metrics) — as INPUTS (values as stored: 6 significant digits), the target = the column medians of the metrics, and the
reference's functions' outputs on them (order, and through the public calls: PRESS, component counts, distances; doubled variance and
weights of the top tenth against the next tenth).
`sampling`: ABC::setup_mvn_sampler's factor (deterministic: exact pin) for the shapes of tests/test_sampling.py, and 10000 draws each
of ABC::sample_predictive_priors / sample_mvn_predictive_priors (src/AbcUtil.cpp:378-404, Priors.h:18-41) on the stand-in's MT19937
stream for that file's two cases — samples of the reference's own rejection / recast / fall-back code, compared distributionally.
Consumers: tests/test_ref_pin.py (oracle vs these, CPU, everywhere) and tests/test_gpu_golden.py (CUDA path vs these, -m gpu).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

# (tag, synth config, scale): the two AbcSmc shapes of BASELINE.json scaled down, both PLS kernel types exercised below
SMALL_CASES = [("C2s", "C2", 0.05), ("C3s", "C3", 0.016)]


def path_outputs(ref, met, par, target, n_pp, theta_old, w_old, dv_old):
    """What particle_ranking_PLS computes on the way (same public calls, AbcUtil.cpp:432-456) + the two follow-up calls."""
    N = met.shape[0]
    order = ref.particle_ranking_PLS(met, par, target, 0.5)
    mean = ref.colwise_mean(met); sd = ref.colwise_stdev(met, mean)
    z_met = ref.colwise_z_scores(met, mean, sd); z_par = ref.colwise_z_scores(par)
    obs = ref.z_scores(target, mean, sd)
    n_tr = int(np.floor(N * 0.5 + 0.5))                   # std::round (AbcUtil.cpp:438)
    m = ref.Model(z_met[:n_tr], z_par[:n_tr])
    em = m.cv_NEW_DATA(z_met[n_tr:], z_par[n_tr:])
    press = em.validation(ref.RESS)
    ncomp = em.optimal_num_components(0.1)
    used = int(ncomp.max())
    dist = ref.euclidean(m.scores(z_met, used), m.scores(obs.reshape(1, -1), used)[0])
    assert np.array_equal(ref.ordered(dist), order)       # the replay reproduces the function's own result
    top = order[:n_pp].astype(np.int64)
    sel = np.asfortranarray(par[top, :])
    dv = ref.calculate_doubled_variance(sel)
    P = par.shape[1]
    res = dict(order=order, mean=mean, sd=sd, press=press, ncomp=ncomp, ncomp_used=used, R=m.R, coef=m.coefficients(used), dist=dist, dv=dv)
    if theta_old is not None:
        res["w"] = ref.weight_predictive_prior([0] * P, [0.0] * P, [1.0] * P, sel, theta_old, w_old, dv_old)    # ContinuousUniformPrior(0, 1)
    return res


def small():
    from abcsmc_b200 import synth
    import oracle.ref as ref
    out = {}
    for tag, name, scale in SMALL_CASES:
        cfg = synth.make_config(name, scale=scale)
        r = path_outputs(ref, cfg["metrics"], cfg["params"], cfg["target"], cfg["N_pp"], cfg["theta_old"], cfg["w_old"], cfg["dv_old"])
        out[f"{tag}_shape"] = np.array([cfg["N"], cfg["K"], cfg["P"], cfg["N_pp"]])
        out[f"{tag}_scale"] = scale
        for k, v in r.items():
            out[f"{tag}_{k}"] = v
        print(tag, cfg["N"], cfg["K"], cfg["P"], "components", r["ncomp_used"], flush=True)
    # the reference's demo inputs: multi-response toy set and the single-response NIR set (M == 1 skips the eigen-solve, pls.cpp:403)
    toy = np.load(os.path.join(HERE, "toy_inputs.npz"))
    for tag, X, Y, A in (("toy", toy["toyX"], toy["toyY"], 5), ("nir", toy["nir"], toy["octane"].reshape(-1, 1), 6)):
        X = ref.colwise_z_scores(X); Y = ref.colwise_z_scores(Y)      # as lib/PLS/src/main.cpp does before fitting
        out[f"{tag}_A"] = A
        for method in (ref.KERNEL_TYPE1, ref.KERNEL_TYPE2):
            m = ref.Model(X, Y, method, A)
            key = f"{tag}_m{method}"
            out[f"{key}_R"] = m.R; out[f"{key}_P"] = m.P; out[f"{key}_W"] = m.W; out[f"{key}_Q"] = m.Q
            out[f"{key}_coef"] = m.coefficients()
            out[f"{key}_sse"] = m.SSE(X, Y); out[f"{key}_ev"] = m.explained_variance(X, Y)
            loo = m.cv_LOO()
            out[f"{key}_loo_press"] = loo.validation(ref.RESS); out[f"{key}_loo_ncomp"] = loo.optimal_num_components(0.1)
            n = X.shape[0]; test_size = int(0.3 * n + 0.5)
            lso = ref.cv_LSO_seeded(m, 0.3, 4, 12345)
            out[f"{key}_lso_press"] = lso.validation(ref.RESS)
            out[f"{key}_lso_shuffles"] = ref.lso_shuffles(12345, n, test_size, 4)
        print(tag, X.shape, Y.shape, flush=True)
    # rank-sum test and the normal CDF approximation on fixed inputs
    rng = np.random.default_rng(20261018)
    e1 = rng.normal(size=4001); e2 = e1 + 0.02 * rng.normal(size=4001) + 0.0005
    out["wilcoxon_e1"] = e1; out["wilcoxon_e2"] = e2
    out["wilcoxon_p"] = np.array([ref.wilcoxon(e1, e2), ref.wilcoxon(e2, e1), ref.wilcoxon(e1, e1)])
    zs = np.linspace(-6, 6, 49)
    out["normalcdf_z"] = zs; out["normalcdf"] = np.array([ref.normalcdf(z) for z in zs])
    # prior likelihoods (Priors.h) and the converged-parameter rule of the weight update (AbcUtil.cpp:573)
    vals = np.array([-0.5, 0.0, 0.25, 1.0, 1.5, 2.0, 3.0, 7.0])
    out["prior_vals"] = vals
    out["prior_lik"] = np.array([[ref.prior_likelihood(t, a, b, v) for v in vals] for t, a, b in ((0, 0.0, 2.0), (1, 0.0, 3.0), (2, 1.0, 0.5))])
    th_old = rng.uniform(size=(60, 3)); th_new = rng.uniform(size=(50, 3)); w_old = rng.uniform(size=60)
    th_old[:, 1] = 0.5; th_new[:, 1] = 0.5                 # a converged parameter: dv == 0 and equal values -> factor skipped
    dv = ref.calculate_doubled_variance(th_old)
    out["wconv_th_old"] = th_old; out["wconv_th_new"] = th_new; out["wconv_w_old"] = w_old; out["wconv_dv"] = dv
    out["wconv_w"] = ref.weight_predictive_prior([0, 0, 2], [0.0, 0.0, 0.5], [1.0, 1.0, 0.3], th_new, th_old, w_old, dv)
    out["w0"] = ref.weight_predictive_prior0(7, 3)
    # filtering-report statistics (AbcUtil.cpp:46-61, 326-345)
    mets = rng.normal(size=(33, 4)) + 2.0; obs = np.array([2.0, 1.5, 0.0, mets[:, 3].mean()])
    out["nrmse_mets"] = mets; out["nrmse_obs"] = obs; out["nrmse"] = ref.calculate_nrmse(mets, obs)
    out["median_in"] = mets[:, 0]; out["median"] = np.array([ref.median(mets[:, 0]), ref.median(mets[:32, 0])])
    path = os.path.join(HERE, "ref_small.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "kB")


def full(name):
    from abcsmc_b200 import synth
    import oracle.ref as ref
    cfg = synth.make_config(name, scale=1.0)
    t0 = time.perf_counter()
    order = ref.particle_ranking_PLS(cfg["metrics"], cfg["params"], cfg["target"], 0.5)
    secs = time.perf_counter() - t0
    path = os.path.join(HERE, f"ref_fullsize_{name}.npz")
    np.savez_compressed(path, name=name, N=cfg["N"], K=cfg["K"], P=cfg["P"], N_pp=cfg["N_pp"], order_top=order[:cfg["N_pp"]],
                        order_checksum=np.uint64(np.bitwise_xor.reduce(order * np.arange(1, order.size + 1, dtype=np.uint64))), seconds=secs)
    print(f"{path}: N={cfg['N']} in {secs:.1f}s", flush=True)


def write_toy_csvs(directory):
    """toyX.csv / toyY.csv as the reference ships them (lib/PLS/toyX.csv, toyY.csv), rewritten from toy_inputs.npz at full precision."""
    toy = np.load(os.path.join(HERE, "toy_inputs.npz"))
    paths = []
    for k in ("toyX", "toyY"):
        path = os.path.join(directory, k + ".csv")
        np.savetxt(path, toy[k].reshape(toy[k].shape[0], -1), delimiter=",", fmt="%.17g")
        paths.append(path)
    return paths


def demo_program():
    import subprocess
    import tempfile
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp"), "_build/pls_main_reference"])
    with tempfile.TemporaryDirectory() as d:
        x, y = write_toy_csvs(d)
        r = subprocess.run([os.path.join(ROOT, "tests", "cpp", "_build", "pls_main_reference"), x, y, "5"], capture_output=True, text=True, check=True)
    path = os.path.join(HERE, "ref_main_toy.txt")
    open(path, "w").write(r.stderr)
    print(path, len(r.stderr), "bytes")


def real_sets():
    """(tag, params N x P, metrics N x K) from the data files under /root/reference, rows in the files' own order."""
    import sqlite3
    import oracle.ref as ref
    con = sqlite3.connect("file:" + os.path.join(ref.REFERENCE_ROOT, "examples", "scratch", "posterior.sqlite") + "?mode=ro", uri=True)
    rows = con.execute("select P.caseEF, P.mos_mov, P.exp_coef, P.num_mos, P.beta, M.mean, M.median, M.stdev, M.max, M.skewness, M.mc, M.sp "
                       "from jobs J, parameters P, metrics M where J.serial = P.serial and J.serial = M.serial order by J.serial;").fetchall()
    con.close()
    a = np.array(rows, dtype=np.float64)
    yield "dengue_sqlite", np.asfortranarray(a[:, :5]), np.asfortranarray(a[:, 5:])
    b = np.loadtxt(os.path.join(ref.REFERENCE_ROOT, "vis", "dengue_predictive_prior-full_ts.06"), skiprows=1)
    yield "dengue_pp250", np.asfortranarray(b[:, 3:7]), np.asfortranarray(b[:, 7:])


def realdata():
    import oracle.ref as ref
    out = {}
    for tag, par, met in real_sets():
        N = par.shape[0]
        target = np.median(met, axis=0)
        n_pp = N // 10
        res = path_outputs(ref, met, par, target, n_pp, None, None, None)
        order = res["order"].astype(np.int64)
        # the weight update on real rows: the top tenth against the next tenth as the "previous" set (uniform weights)
        th_new, th_old = np.asfortranarray(par[order[:n_pp]]), np.asfortranarray(par[order[n_pp:2 * n_pp]])
        dv_old = ref.calculate_doubled_variance(th_old)
        w_old = np.full(n_pp, 1.0 / n_pp)
        lo, hi = par.min(axis=0) - 1.0, par.max(axis=0) + 1.0
        w = ref.weight_predictive_prior([0] * par.shape[1], lo, hi, th_new, th_old, w_old, dv_old)
        out.update({f"{tag}_par": par, f"{tag}_met": met, f"{tag}_target": target, f"{tag}_prior_lo": lo, f"{tag}_prior_hi": hi, f"{tag}_w_vs_next": w, f"{tag}_dv_next": dv_old})
        out.update({f"{tag}_{k}": v for k, v in res.items()})
        print(tag, par.shape, met.shape, "ncomp", res["ncomp"], "ties in dist:", int(np.sum(np.diff(np.sort(res["dist"])) == 0)))
    path = os.path.join(HERE, "ref_realdata.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


def mvn_shapes():
    """The inputs of tests/test_sampling.py::test_setup_mvn_sampler_matches_oracle, regenerated from their seeds."""
    for shape, seed in (((600, 3), 9), ((5000, 30), 1), ((77, 10), 2)):
        r = np.random.default_rng(seed)
        yield shape, seed, np.asfortranarray(r.normal(size=shape) @ r.normal(size=(shape[1], shape[1])) * 0.1 + r.random(shape[1]))


def sampling():
    import oracle.ref as ref
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_sampling import _case, _mvn_case
    out = {}
    for shape, seed, th in mvn_shapes():
        out[f"L_{shape[0]}x{shape[1]}_s{seed}"] = ref.setup_mvn_sampler(th)
    n = 10000
    c = _case()
    out["indep_samples"] = ref.sample_predictive_priors(20261018, n, c["w"], c["theta"], c["ptype"], c["pa"], c["pb"], c["dv"])
    m = _mvn_case()
    L = ref.setup_mvn_sampler(m["theta"])
    out["mvn_L"] = L
    out["mvn_samples"] = ref.sample_mvn_predictive_priors(20261019, n, m["w"], m["theta"], m["ptype"], m["pa"], m["pb"], L)
    # the fall-back branches: parents far outside U(0, 1) -> the prior's mean after MAX_ATTEMPTS (Priors.h:26-28)
    far = np.asfortranarray(np.full((5, 1), 50.0))
    out["indep_fallback"] = ref.sample_predictive_priors(1, 8, np.ones(5), far, [0], [0.0], [1.0], [1e-4])
    path = os.path.join(HERE, "ref_sampling.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "small"
    if what == "small":
        small()
    elif what == "realdata":
        realdata()
    elif what == "sampling":
        sampling()
    elif what == "main":
        demo_program()
    else:
        full(what)
