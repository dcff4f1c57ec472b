"""The C++ host adapter (abcsmc_b200/host/abc_b200.hpp): the reference's own signatures on top of the C ABI.
CPU: the header compiles (-Wall -Wpedantic) against a minimal Eigen-like matrix type and links to libabcsmc_b200.so.
GPU: the AbcSmc call sequence (rank -> truncate -> gather -> doubled variance -> weights) through the adapter matches the
CPU oracle (bit-exact indices, 1e-10 on FP64 outputs)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from abcsmc_b200 import _capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "adapter_test.cpp")


def _build(tmp_path):
    _capi.build()
    exe = str(tmp_path / "adapter_test")
    libdir = os.path.join(ROOT, "abcsmc_b200")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wpedantic", "-Werror", "-O1", SRC, f"-L{libdir}", "-labcsmc_b200",
                           f"-Wl,-rpath,{libdir}", "-o", exe])
    return exe


def test_adapter_compiles_and_links(tmp_path):
    exe = _build(tmp_path)
    assert os.path.exists(exe)
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-x", "c++", os.path.join(ROOT, "abcsmc_b200", "host", "abc_b200.hpp")])


def _build_dropin(tmp_path):
    _capi.build()
    exe = str(tmp_path / "dropin_test")
    libdir = os.path.join(ROOT, "abcsmc_b200")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wpedantic", "-Werror", "-O1", os.path.join(ROOT, "tests", "cpp", "dropin_test.cpp"),
                           f"-L{libdir}", "-labcsmc_b200", f"-Wl,-rpath,{libdir}", "-o", exe])
    return exe


def test_dropin_blocks_compile_and_link(tmp_path):
    """The ABCB200_DROP_IN (+ _SAMPLER, + _PLS) blocks INTEGRATION.md tells a maintainer to use, compiled against stand-ins of the
    reference's typedefs (tests/cpp/ref_stub.hpp) and linked with the library's C ABI."""
    assert os.path.exists(_build_dropin(tmp_path))


REFHDR_EXE = os.path.join(ROOT, "tests", "cpp", "_build", "dropin_refhdr_test")
REFERENCE_ROOT = os.environ.get("ABC_REFERENCE_ROOT", "/root/reference")


def test_dropin_compiles_against_reference_headers():
    """INTEGRATION.md §2 with the reference's REAL headers: <AbcSmc/AbcUtil.h> + <AbcSmc/Priors.h> from /root/reference (unmodified;
    Eigen / GSL through the stand-ins of oracle/shim/), then abc_b200.hpp with ABCB200_DROP_IN. The program calls the five ABC::
    functions through the reference's own declarations and links WITHOUT src/AbcUtil.cpp, so it only links if the drop-in's
    signatures are the reference's. Only where the reference's sources exist (not on the GPU box)."""
    if not os.path.exists(os.path.join(REFERENCE_ROOT, "include", "AbcSmc", "AbcUtil.h")):
        pytest.skip("reference headers not present")
    _capi.build()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp"), "_build/dropin_refhdr_test", "REF=" + REFERENCE_ROOT])
    assert os.path.exists(REFHDR_EXE)


PLS_MAIN_EXE = os.path.join(ROOT, "tests", "cpp", "_build", "pls_main_dropin")


def test_reference_demo_program_compiles_on_the_pls_dropin():
    """The reference's lib/PLS/src/main.cpp, UNMODIFIED, compiled with tests/cpp/pls_dropin_inc/PLS/pls.h (typedefs + abc_b200.hpp's
    ABCB200_DROP_IN_PLS block) in place of the reference's header and linked with libabcsmc_b200.so: every name the program uses
    (read_matrix_file, colwise_z_scores, Model, METHOD::KERNEL_TYPE1, print_state, print_explained_variance, cv_LOO, cv_LSO with a
    std::mt19937, Residual, print_validation, MSE) resolves to the CUDA-backed namespace PLS. SURVEY §8 row a13."""
    if not os.path.exists(os.path.join(REFERENCE_ROOT, "lib", "PLS", "src", "main.cpp")):
        pytest.skip("reference sources not present")
    _capi.build()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp"), "_build/pls_main_dropin", "REF=" + REFERENCE_ROOT])
    assert os.path.exists(PLS_MAIN_EXE)


def _tokens(text, complex_pairs=False):
    import re
    if complex_pairs:                                     # the reference streams complex factors as (re,im): keep the real part
        text = re.sub(r"\(([^,()]+),[^()]*\)", r"\1", text)
    return re.findall(r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?|nan|inf|[A-Za-z_#()=;:]+", text)


def test_flatten_prior_host_logic(tmp_path):
    """No GPU: the adapter's Parameter -> (lo, hi, integral, mean) flattening against the three prior kinds of Priors.h."""
    exe = str(tmp_path / "flatten_test")
    libdir = os.path.join(ROOT, "abcsmc_b200")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wpedantic", "-Werror", "-O1", os.path.join(ROOT, "tests", "cpp", "flatten_test.cpp"),
                           f"-L{libdir}", "-labcsmc_b200", f"-Wl,-rpath,{libdir}", "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "flatten ok" in out.stdout, out.stderr


@pytest.mark.gpu
def test_adapter_matches_oracle(tmp_path, oracle):
    exe = _build(tmp_path)
    cfg = synth.make_config("C2", scale=0.05)
    N, K, P, Npp = cfg["N"], cfg["K"], cfg["P"], cfg["N_pp"]
    th_old, w_old, dv_old = cfg["theta_old"], cfg["w_old"], cfg["dv_old"]
    case, out = tmp_path / "case.bin", tmp_path / "out.bin"
    with open(case, "wb") as f:
        f.write(struct.pack("5q", N, K, P, Npp, th_old.shape[0]))
        for a in (cfg["metrics"], cfg["params"], cfg["target"], th_old, w_old, dv_old):
            f.write(np.asfortranarray(a, dtype=np.float64).tobytes(order="F"))
    subprocess.check_call([exe, str(case), str(out)])
    raw = open(out, "rb").read()
    off = 0

    def take(dtype, n):
        nonlocal off
        a = np.frombuffer(raw, dtype=dtype, count=n, offset=off); off += a.nbytes
        return a
    order, dv, w0, w, simple, B = take(np.int64, Npp), take(np.float64, P), take(np.float64, Npp), take(np.float64, Npp), take(np.int64, Npp), take(np.float64, K * P)
    prop = take(np.float64, 2 * Npp * P).reshape(P, 2 * Npp).T
    Nl = min(N, 300)
    nloo, press_loo = take(np.int64, P), take(np.float64, P * K).reshape(K, P).T
    nlso, press_lso = take(np.int64, P), take(np.float64, P * K).reshape(K, P).T
    shuf = take(np.int64, 3 * Nl).reshape(3, Nl)

    o = oracle.particle_ranking_PLS(cfg["metrics"], cfg["params"], cfg["target"], 0.5)
    assert np.array_equal(order, o["order"][:Npp].astype(np.int64))
    sel = cfg["params"][order, :]
    np.testing.assert_allclose(dv, oracle.calculate_doubled_variance(sel), rtol=1e-10)
    assert np.all(w0 == 1.0 / Npp)
    numer = np.full(Npp, 0.5 ** P)                       # uniform prior on [0, 2] in every dimension
    np.testing.assert_allclose(w, oracle.weight_predictive_prior(numer, sel, th_old, w_old, dv_old), rtol=1e-10)
    assert np.array_equal(simple, oracle.particle_ranking_simple(cfg["metrics"], cfg["target"])["order"][:Npp].astype(np.int64))
    # Model::cv_LOO / cv_LSO through the C++ wrapper against the oracle's Residual (same partitions for LSO)
    om = oracle.Model(cfg["metrics"][:Nl], cfg["params"][:Nl], 0)
    r = om.cv_LOO()
    np.testing.assert_allclose(press_loo, r.validation(oracle.RESS), rtol=1e-10)
    assert list(nloo) == [int(v) for v in r.optimal_num_components(0.1)]
    r = om.cv_LSO(shuf, int(0.2 * Nl + 0.5))
    np.testing.assert_allclose(press_lso, r.validation(oracle.RESS), rtol=1e-10)
    assert list(nlso) == [int(v) for v in r.optimal_num_components(0.1)]
    # proposals: inside the prior's support, centred on the weighted predictive prior with the doubled variance added
    assert prop.shape == (2 * Npp, P) and prop.min() >= 0.0 and prop.max() <= 2.0
    wn = w / w.sum()
    mu = wn @ sel
    var = wn @ (sel - mu) ** 2 + dv
    assert np.all(np.abs(prop.mean(axis=0) - mu) < 6 * np.sqrt(var / (2 * Npp)) + 1e-3)
    Bo = oracle.Model(cfg["metrics"], cfg["params"], 0).coefficients()
    np.testing.assert_allclose(B.reshape(P, K).T, Bo, rtol=0, atol=1e-10 * np.abs(Bo).max())


@pytest.mark.gpu
def test_dropin_matches_oracle(tmp_path, oracle):
    """namespace ABC as AbcSmc.cpp drives it and namespace PLS as lib/PLS/src/main.cpp drives it, through the drop-in blocks."""
    exe = _build_dropin(tmp_path)
    cfg = synth.make_config("C2", scale=0.03)
    N, K, P, Npp = cfg["N"], cfg["K"], cfg["P"], cfg["N_pp"]
    th_old, w_old, dv_old = cfg["theta_old"], cfg["w_old"], cfg["dv_old"]
    case, out = tmp_path / "case.bin", tmp_path / "out.bin"
    with open(case, "wb") as f:
        f.write(struct.pack("5q", N, K, P, Npp, th_old.shape[0]))
        for a in (cfg["metrics"], cfg["params"], cfg["target"], th_old, w_old, dv_old):
            f.write(np.asfortranarray(a, dtype=np.float64).tobytes(order="F"))
    subprocess.check_call([exe, str(case), str(out)])
    raw = open(out, "rb").read()
    off = 0

    def take(dtype, n):
        nonlocal off
        a = np.frombuffer(raw, dtype=dtype, count=n, offset=off); off += a.nbytes
        return a
    order, dv, w0, w, simple = take(np.int64, Npp), take(np.float64, P), take(np.float64, Npp), take(np.float64, Npp), take(np.int64, Npp)
    dist, prop = take(np.float64, Npp), take(np.float64, Npp * P).reshape(P, Npp).T
    Nl, Nh, A, n_lso = (int(v) for v in take(np.int64, 4))
    ev_loo = take(np.float64, P * Nl * A).reshape(P, A, Nl)
    mse_loo, nc_loo = take(np.float64, P * A).reshape(A, P).T, take(np.uint64, P)
    ev_nd = take(np.float64, P * Nh * A).reshape(P, A, Nh)
    press_nd, nc_nd, nc_nd05 = take(np.float64, P * A).reshape(A, P).T, take(np.uint64, P), take(np.uint64, P)
    expl = take(np.float64, P)
    assert off == len(raw)

    o = oracle.particle_ranking_PLS(cfg["metrics"], cfg["params"], cfg["target"], 0.5)
    assert np.array_equal(order, o["order"][:Npp].astype(np.int64))
    sel = cfg["params"][order, :]
    np.testing.assert_allclose(dv, oracle.calculate_doubled_variance(sel), rtol=1e-10)
    assert np.all(w0 == 1.0 / Npp)
    np.testing.assert_allclose(w, oracle.weight_predictive_prior(np.full(Npp, 0.5 ** P), sel, th_old, w_old, dv_old), rtol=1e-10)
    assert np.array_equal(simple, oracle.particle_ranking_simple(cfg["metrics"], cfg["target"])["order"][:Npp].astype(np.int64))
    np.testing.assert_allclose(dist, oracle.euclidean(sel, dv), rtol=1e-12)
    assert prop.min() >= 0.0 and prop.max() <= 2.0

    X = oracle.colwise_z_scores(cfg["metrics"][:Nl]); Y = oracle.colwise_z_scores(cfg["params"][:Nl])
    om = oracle.Model(X, Y, 0, A)
    r = om.cv_LOO()
    for y, e in enumerate(r.errors()):
        np.testing.assert_allclose(ev_loo[y].T, e, rtol=0, atol=1e-10 * np.abs(e).max())
    np.testing.assert_allclose(mse_loo, r.validation(oracle.MSE), rtol=1e-10)
    assert list(nc_loo) == [int(v) for v in r.optimal_num_components(0.1)]
    r = om.cv_NEW_DATA(cfg["metrics"][Nl:Nl + Nh], cfg["params"][Nl:Nl + Nh])
    for y, e in enumerate(r.errors()):
        np.testing.assert_allclose(ev_nd[y].T, e, rtol=0, atol=1e-10 * np.abs(e).max())
    np.testing.assert_allclose(press_nd, r.validation(oracle.RESS), rtol=1e-10)
    assert list(nc_nd) == [int(v) for v in r.optimal_num_components(0.1)]
    assert list(nc_nd05) == [int(v) for v in r.optimal_num_components(0.05)]
    np.testing.assert_allclose(expl, om.explained_variance(X, Y), rtol=1e-10)


def _duplicate_leading_particles(cfg, oracle, n_src=8, n_dup=20):
    """Copies of particles that rank near the top written over far-ranked hold-out rows: exact distance ties inside the top N_pp."""
    met, par = cfg["metrics"].copy(), cfg["params"].copy()
    base = oracle.particle_ranking_PLS(met, par, cfg["target"], 0.5)["order"].astype(np.int64)
    n_tr = int(round(cfg["N"] * 0.5))
    src = [int(i) for i in base[5:cfg["N_pp"]] if i >= n_tr][:n_src]
    dst = [int(i) for i in base[-1000:] if i >= n_tr][:n_dup]
    for j, d in enumerate(dst):
        met[d, :] = met[src[j % len(src)], :]; par[d, :] = par[src[j % len(src)], :]
    return dict(cfg, metrics=np.asfortranarray(met), params=np.asfortranarray(par))


@pytest.mark.gpu
@pytest.mark.parametrize("tied", [False, True])
def test_dropin_on_reference_headers_matches_oracle(tmp_path, oracle, tied):
    """The executable built from the reference's real headers + the drop-in block (test_dropin_compiles_against_reference_headers;
    prebuilt in the authoring container, it travels in tests/cpp/_build/) run on the GPU: the reference's call sequence with the
    reference's own types and prior classes, every ABC:: call served by the CUDA library, against the oracle."""
    if not os.path.exists(REFHDR_EXE):
        pytest.skip("tests/cpp/_build/dropin_refhdr_test not built (needs the reference's headers: make -C tests/cpp)")
    cfg = synth.make_config("C2", scale=0.05)
    if tied:   # the drop-in places exact ties as libstdc++'s std::sort does in PLS::ordered (abcb200_set_tie_order 1 in the adapter's context)
        cfg = _duplicate_leading_particles(cfg, oracle)
    N, K, P, Npp = cfg["N"], cfg["K"], cfg["P"], cfg["N_pp"]
    th_old, w_old, dv_old = cfg["theta_old"], cfg["w_old"], cfg["dv_old"]
    case, out = tmp_path / "case.bin", tmp_path / "out.bin"
    with open(case, "wb") as f:
        f.write(struct.pack("5q", N, K, P, Npp, th_old.shape[0]))
        for a in (cfg["metrics"], cfg["params"], cfg["target"], th_old, w_old, dv_old):
            f.write(np.asfortranarray(a, dtype=np.float64).tobytes(order="F"))
    subprocess.check_call([REFHDR_EXE, str(case), str(out)])
    raw = open(out, "rb").read()
    off = 0

    def take(dtype, n):
        nonlocal off
        a = np.frombuffer(raw, dtype=dtype, count=n, offset=off); off += a.nbytes
        return a
    order, dv, w0, w, simple, dist = take(np.int64, Npp), take(np.float64, P), take(np.float64, Npp), take(np.float64, Npp), take(np.int64, Npp), take(np.float64, Npp)
    ref = oracle.particle_ranking_PLS(cfg["metrics"], cfg["params"], cfg["target"], 0.5)
    assert np.array_equal(order, ref["order"][:Npp].astype(np.int64))
    if tied:
        assert np.sum(np.diff(ref["dist"][order]) == 0) >= 8                  # the tie groups really are inside what was compared
    sel = np.asfortranarray(cfg["params"][order, :])
    np.testing.assert_allclose(dv, oracle.calculate_doubled_variance(sel), rtol=1e-10)
    assert np.all(w0 == 1.0 / Npp)
    numer = np.full(Npp, 0.5 ** P)                                # ContinuousUniformPrior(0, 2) for every parameter (Priors.h:101-103)
    np.testing.assert_allclose(w, oracle.weight_predictive_prior(numer, sel, th_old, w_old, dv_old), rtol=1e-10)
    assert np.array_equal(simple, oracle.particle_ranking_simple(cfg["metrics"], cfg["target"])["order"][:Npp].astype(np.int64))
    np.testing.assert_allclose(dist, oracle.euclidean(sel, dv), rtol=1e-10)


@pytest.mark.gpu
def test_reference_demo_program_on_gpu_prints_what_the_reference_prints(tmp_path):
    """lib/PLS/src/main.cpp (unmodified, built on the PLS drop-in: tests/cpp/_build/pls_main_dropin, prebuilt where the reference's
    sources exist) run on the toy demo inputs with 5 components, against what the same program prints when built on the reference's
    own pls.cpp (tests/golden/ref_main_toy.txt, make_ref_fixtures.py main): same words, numbers to the 6 digits the streams print;
    factor matrices up to the sign of each component; cv_LSO draws 100 partitions from a default-seeded std::mt19937 in both."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_ref_fixtures import write_toy_csvs
    if not os.path.exists(PLS_MAIN_EXE):
        pytest.skip("tests/cpp/_build/pls_main_dropin not built (needs the reference's sources: make -C tests/cpp)")
    x, y = write_toy_csvs(str(tmp_path))
    r = subprocess.run([PLS_MAIN_EXE, x, y, "5"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    got = _tokens(r.stderr)
    want = _tokens(open(os.path.join(ROOT, "tests", "golden", "ref_main_toy.txt")).read(), complex_pairs=True)
    assert len(got) == len(want), (len(got), len(want))
    in_factors = False
    for g, w in zip(got, want):
        try:
            wv = float(w)
        except ValueError:
            assert g == w, (g, w)
            if w == "P:":
                in_factors = True
            if w == "coefficients:":
                in_factors = False
            continue
        gv = float(g)
        if in_factors:
            gv, wv = abs(gv), abs(wv)
        assert abs(gv - wv) <= 2e-5 * max(abs(wv), 1e-2), (g, w)
