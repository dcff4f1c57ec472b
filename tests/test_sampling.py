"""Next-set proposal sampling: ABC::sample_predictive_priors, src/AbcUtil.cpp:378-390.

The reference consumes a gsl_rng stream, so parity is DISTRIBUTIONAL: the oracle restatement (own splitmix64 stream, the
reference's order of draws) is first checked against closed forms on the CPU (scipy truncated normal, weight frequencies),
then the CUDA path (Philox counters) is checked against the oracle's samples with two-sample tests. Thresholds are set for
fixed seeds at significance 1e-4 or looser, so the tests are deterministic."""
import numpy as np
import pytest
from scipy import stats

PRIOR_UNIFORM, PRIOR_DISCRETE_UNIFORM, PRIOR_GAUSSIAN = 0, 1, 2


def _case(n_pp=400, seed=3):
    """A predictive prior near the lower bound of parameter 0 (truncation matters), an integer parameter and an unbounded one."""
    r = np.random.default_rng(seed)
    theta = np.asfortranarray(np.column_stack([np.abs(r.normal(0.05, 0.05, n_pp)).clip(0, 1),       # U(0, 1), mass at the edge
                                               r.integers(2, 9, n_pp).astype(float),                # discrete U{0..10}
                                               r.normal(1.0, 0.3, n_pp)]))                          # N(1, 2) prior: no truncation
    w = r.random(n_pp) ** 3
    w[::7] = 0.0                                                                                   # zero-weight rows are never drawn
    w /= np.linalg.norm(w)                                                                         # AbcSmc stores L2-normalised weights
    dv = np.array([0.02, 3.0, 0.18])
    ptype = np.array([PRIOR_UNIFORM, PRIOR_DISCRETE_UNIFORM, PRIOR_GAUSSIAN])
    pa, pb = np.array([0.0, 0.0, 1.0]), np.array([1.0, 10.0, 2.0])
    lo, hi = np.array([0.0, 0.0, -np.inf]), np.array([1.0, 10.0, np.inf])
    mean = np.array([0.5, 5.0, 1.0])
    integral = np.array([0, 1, 0], dtype=np.int32)
    return dict(theta=theta, w=w, dv=dv, ptype=ptype, pa=pa, pb=pb, lo=lo, hi=hi, mean=mean, integral=integral)


# ---- the oracle against closed forms (CPU) -------------------------------------------------------------------------------
def test_oracle_sampler_matches_closed_forms(oracle):
    c = _case()
    n = 60000
    r = oracle.sample_predictive_priors(11, n, c["w"], c["theta"], c["ptype"], c["pa"], c["pb"], c["dv"])
    s, parent = r["samples"], r["parent"].astype(np.int64)
    assert r["fallbacks"] == 0
    # rows are drawn with P(j) = w_j / sum w (gsl_ran_discrete, AbcUtil.cpp:111-121)
    counts = np.bincount(parent, minlength=c["w"].size)
    assert counts[c["w"] == 0].sum() == 0
    live = c["w"] > 0
    chi2 = stats.chisquare(counts[live], n * c["w"][live] / c["w"].sum())
    assert chi2.pvalue > 1e-4
    # parameter 0: truncated normal around the parent (rejection until inside [0, 1], Priors.h:18-33): probability integral transform
    mu, sd = c["theta"][parent, 0], np.sqrt(c["dv"][0])
    a, b = (0.0 - mu) / sd, (1.0 - mu) / sd
    u = (stats.norm.cdf((s[:, 0] - mu) / sd) - stats.norm.cdf(a)) / (stats.norm.cdf(b) - stats.norm.cdf(a))
    assert stats.kstest(u, "uniform").pvalue > 1e-4
    assert s[:, 0].min() >= 0.0 and s[:, 0].max() <= 1.0
    # parameter 1: integers in range (DiscreteUniformPrior::recast rounds, Priors.h:80)
    assert np.all(s[:, 1] == np.round(s[:, 1])) and s[:, 1].min() >= 0 and s[:, 1].max() <= 10
    # parameter 2: Gaussian prior, never invalid: plain normal noise
    z = (s[:, 2] - c["theta"][parent, 2]) / np.sqrt(c["dv"][2])
    assert stats.kstest(z, "norm").pvalue > 1e-4


def test_oracle_sampler_falls_back_to_prior_mean(oracle):
    """Priors.h:26-28: after MAX_ATTEMPTS invalid draws the prior's mean is returned."""
    theta = np.asfortranarray(np.full((5, 1), 50.0))           # parents far outside U(0, 1): every draw is invalid
    r = oracle.sample_predictive_priors(1, 40, np.ones(5), theta, [PRIOR_UNIFORM], [0.0], [1.0], [1e-4], max_attempts=20)
    assert r["fallbacks"] == 40 and np.all(r["samples"] == 0.5)


# ---- the CUDA path against the oracle (GPU) --------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def api():
    from abcsmc_b200 import api as a
    a.get_context(0)
    return a


@pytest.mark.gpu
def test_sample_predictive_priors_distribution(api, oracle):
    c = _case()
    n = 60000
    g = api.sample_predictive_priors(12345, n, c["w"], c["theta"], c["dv"], c["lo"], c["hi"], c["mean"], integral=c["integral"], return_info=True)
    o = oracle.sample_predictive_priors(999, n, c["w"], c["theta"], c["ptype"], c["pa"], c["pb"], c["dv"])
    s, parent = g["samples"], g["parent"].astype(np.int64)
    assert g["fallbacks"] == 0 and s.shape == (n, 3)
    # parent rows: exact target frequencies, and the same distribution as the oracle's draw
    counts = np.bincount(parent, minlength=c["w"].size)
    assert counts[c["w"] == 0].sum() == 0
    live = c["w"] > 0
    assert stats.chisquare(counts[live], n * c["w"][live] / c["w"].sum()).pvalue > 1e-4
    # every output column: two-sample KS against the oracle's samples
    for p in range(3):
        assert stats.ks_2samp(s[:, p], o["samples"][:, p]).pvalue > 1e-4, p
    # support and recast
    assert s[:, 0].min() >= 0.0 and s[:, 0].max() <= 1.0
    assert np.all(s[:, 1] == np.round(s[:, 1])) and s[:, 1].min() >= 0 and s[:, 1].max() <= 10
    # noise around the parent: probability integral transform of the truncated normal (parameter 0), plain normal (parameter 2)
    mu, sd = c["theta"][parent, 0], np.sqrt(c["dv"][0])
    a, b = (0.0 - mu) / sd, (1.0 - mu) / sd
    u = (stats.norm.cdf((s[:, 0] - mu) / sd) - stats.norm.cdf(a)) / (stats.norm.cdf(b) - stats.norm.cdf(a))
    assert stats.kstest(u, "uniform").pvalue > 1e-4
    assert stats.kstest((s[:, 2] - c["theta"][parent, 2]) / np.sqrt(c["dv"][2]), "norm").pvalue > 1e-4
    # parameters of one sample are noised independently (AbcUtil.cpp:152-156)
    z0 = s[:, 0] - c["theta"][parent, 0]; z2 = s[:, 2] - c["theta"][parent, 2]
    assert abs(np.corrcoef(z0, z2)[0, 1]) < 0.02


@pytest.mark.gpu
def test_sample_predictive_priors_reproducible_and_seeded(api):
    c = _case(n_pp=97)
    a1 = api.sample_predictive_priors(7, 1000, c["w"], c["theta"], c["dv"], c["lo"], c["hi"], c["mean"], integral=c["integral"])
    a2 = api.sample_predictive_priors(7, 1000, c["w"], c["theta"], c["dv"], c["lo"], c["hi"], c["mean"], integral=c["integral"])
    a3 = api.sample_predictive_priors(8, 1000, c["w"], c["theta"], c["dv"], c["lo"], c["hi"], c["mean"], integral=c["integral"])
    assert np.array_equal(a1, a2) and not np.array_equal(a1, a3)
    # a prefix of a longer draw is the shorter draw: every (sample, parameter, attempt) owns its counter
    a4 = api.sample_predictive_priors(7, 1500, c["w"], c["theta"], c["dv"], c["lo"], c["hi"], c["mean"], integral=c["integral"])
    assert np.array_equal(a4[:1000], a1)


@pytest.mark.gpu
def test_sample_predictive_priors_edges(api):
    # fall-back to the prior mean after max_attempts invalid draws (Priors.h:26-28), counted
    theta = np.asfortranarray(np.full((5, 1), 50.0))
    r = api.sample_predictive_priors(1, 40, np.ones(5), theta, [1e-4], [0.0], [1.0], [0.5], max_attempts=20, return_info=True)
    assert r["fallbacks"] == 40 and np.all(r["samples"] == 0.5)
    # zero variance (calculate_doubled_variance of identical rows): the parent's value itself
    theta = np.asfortranarray(np.array([[0.25], [0.75]]))
    r = api.sample_predictive_priors(2, 200, [1.0, 3.0], theta, [0.0], [0.0], [1.0], [0.5], return_info=True)
    assert np.array_equal(r["samples"][:, 0], theta[r["parent"].astype(np.int64), 0]) and r["fallbacks"] == 0
    assert 0.6 < np.mean(r["parent"] == 1) < 0.9
    # a single row, a single sample
    r = api.sample_predictive_priors(3, 1, [2.0], np.asfortranarray([[0.5, 4.0]]), [0.01, 1.0], [0.0, 0.0], [1.0, 10.0], [0.5, 5.0],
                                     integral=[0, 1], return_info=True)
    assert r["samples"].shape == (1, 2) and r["parent"][0] == 0
    # invalid weight tables are rejected like gsl_ran_discrete_preproc does (abort there, error code here)
    from abcsmc_b200._capi import Abcb200Error
    for bad in ([0.0, 0.0], [1.0, -1.0], [1.0, np.nan]):
        with pytest.raises(Abcb200Error):
            api.sample_predictive_priors(4, 10, bad, np.asfortranarray([[0.1], [0.2]]), [0.01], [0.0], [1.0], [0.5])


# ---- multivariate noise (NOISE::MULTIVARIATE): setup_mvn_sampler + sample_mvn_predictive_priors -------------------------
def _mvn_case(n_pp=600, seed=9):
    r = np.random.default_rng(seed)
    A = np.array([[0.08, 0.0, 0.0], [0.05, 0.06, 0.0], [-0.2, 0.1, 0.3]])
    theta = np.asfortranarray(np.array([0.5, 0.5, 1.0]) + r.normal(size=(n_pp, 3)) @ A.T)    # correlated predictive prior
    theta[:, 0] = np.clip(theta[:, 0], 0.0, 1.0); theta[:, 1] = np.clip(theta[:, 1], 0.0, 1.0)
    w = r.random(n_pp) + 0.1
    ptype = np.array([PRIOR_UNIFORM, PRIOR_UNIFORM, PRIOR_GAUSSIAN])
    pa, pb = np.array([0.0, 0.0, 1.0]), np.array([1.0, 1.0, 2.0])
    lo, hi = np.array([0.0, 0.0, -np.inf]), np.array([1.0, 1.0, np.inf])
    return dict(theta=theta, w=w, ptype=ptype, pa=pa, pb=pb, lo=lo, hi=hi)


def test_oracle_mvn_sampler(oracle):
    c = _mvn_case()
    L = oracle.setup_mvn_sampler(c["theta"])
    S = np.cov(c["theta"], rowvar=False, ddof=1)
    S[np.diag_indices(3)] *= 2.0                                               # AbcUtil.cpp:477-480
    np.testing.assert_allclose(L @ L.T, S, rtol=1e-12, atol=1e-15)
    assert np.allclose(np.triu(L, 1), 0.0)
    # unbounded priors: the noise is exactly N(0, L L^T)
    pt = np.full(3, PRIOR_GAUSSIAN)
    r = oracle.sample_mvn_predictive_priors(5, 40000, c["w"], c["theta"], pt, np.zeros(3), np.full(3, 1e3), L)
    dz = r["samples"] - c["theta"][r["parent"].astype(np.int64)]
    np.testing.assert_allclose(np.cov(dz, rowvar=False), S, rtol=0, atol=4 * np.abs(S).max() / np.sqrt(40000) * 3)
    assert r["failures"] == 0


@pytest.mark.gpu
def test_setup_mvn_sampler_matches_oracle(api, oracle):
    for shape, seed in (((600, 3), 9), ((5000, 30), 1), ((77, 10), 2)):
        r = np.random.default_rng(seed)
        th = np.asfortranarray(r.normal(size=shape) @ r.normal(size=(shape[1], shape[1])) * 0.1 + r.random(shape[1]))
        np.testing.assert_allclose(api.setup_mvn_sampler(th), oracle.setup_mvn_sampler(th), rtol=1e-10, atol=1e-13)
    # a constant column makes the matrix singular: error code here, GSL_EDOM in the reference
    from abcsmc_b200._capi import Abcb200Error
    bad = np.asfortranarray(np.column_stack([np.arange(50.0), np.ones(50)]))
    with pytest.raises(Abcb200Error):
        api.setup_mvn_sampler(bad)


@pytest.mark.gpu
def test_sample_mvn_predictive_priors_distribution(api, oracle):
    c = _mvn_case()
    n = 50000
    L = api.setup_mvn_sampler(c["theta"])
    g = api.sample_mvn_predictive_priors(4242, n, c["w"], c["theta"], L, c["lo"], c["hi"], return_info=True)
    o = oracle.sample_mvn_predictive_priors(77, n, c["w"], c["theta"], c["ptype"], c["pa"], c["pb"], L)
    s, parent = g["samples"], g["parent"].astype(np.int64)
    assert g["failures"] == 0 and o["failures"] == 0
    assert s[:, :2].min() >= 0.0 and s[:, :2].max() <= 1.0
    counts = np.bincount(parent, minlength=c["w"].size)
    assert stats.chisquare(counts, n * c["w"] / c["w"].sum()).pvalue > 1e-4
    for p in range(3):
        assert stats.ks_2samp(s[:, p], o["samples"][:, p]).pvalue > 1e-4, p
    # joint structure: covariance of the accepted noise agrees with the oracle's (truncation included)
    dg = s - c["theta"][parent]; do = o["samples"] - c["theta"][o["parent"].astype(np.int64)]
    Cg, Co = np.cov(dg, rowvar=False), np.cov(do, rowvar=False)
    assert np.all(np.abs(Cg - Co) < 0.05 * np.sqrt(np.outer(np.diag(Co), np.diag(Co))))
    assert abs(np.corrcoef(dg[:, 0], dg[:, 1])[0, 1]) > 0.3                     # the off-diagonal of L really is applied
    # reproducible, and a longer draw extends a shorter one
    g2 = api.sample_mvn_predictive_priors(4242, 1000, c["w"], c["theta"], L, c["lo"], c["hi"])
    assert np.array_equal(g2, s[:1000])
    # an impossible support runs out of attempts: the recast parent row, counted
    r = api.sample_mvn_predictive_priors(1, 64, c["w"], c["theta"], L, [5.0, 5.0, 5.0], [6.0, 6.0, 6.0], max_attempts=10, return_info=True)
    assert r["failures"] == 64 and np.array_equal(r["samples"], c["theta"][r["parent"].astype(np.int64)])


# ---- pinned to samples of the reference's OWN sampling code (tests/golden/ref_sampling.npz, make_ref_fixtures.py sampling) -------
# ABC::sample_predictive_priors / sample_mvn_predictive_priors / setup_mvn_sampler compiled unmodified against the GSL stand-in
# (oracle/shim/gsl/gsl_stub.h: MT19937, polar Box-Muller, cumulative-weight discrete draw, vcov + Cholesky): the rejection, recast
# and fall-back statements are the reference's. The factor is deterministic (exact pin); the samples are compared distributionally.
@pytest.fixture(scope="module")
def ref_fx():
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "ref_sampling.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/ref_sampling.npz not generated (tests/golden/make_ref_fixtures.py sampling)")
    return np.load(path)


def _mvn_shapes():
    for shape, seed in (((600, 3), 9), ((5000, 30), 1), ((77, 10), 2)):
        r = np.random.default_rng(seed)
        yield f"L_{shape[0]}x{shape[1]}_s{seed}", np.asfortranarray(r.normal(size=shape) @ r.normal(size=(shape[1], shape[1])) * 0.1 + r.random(shape[1]))


def _check_against_reference_samples(s, ref_s, theta, parent=None):
    for p in range(s.shape[1]):
        assert stats.ks_2samp(s[:, p], ref_s[:, p]).pvalue > 1e-4, p
    # second moments of the proposals themselves (parents + noise): the joint structure, whatever stream drew it
    Cs, Cr = np.cov(s, rowvar=False), np.cov(ref_s, rowvar=False)
    assert np.all(np.abs(Cs - Cr) < 0.06 * np.sqrt(np.outer(np.diag(Cr), np.diag(Cr))))


def test_oracle_sampling_matches_reference_fixture(oracle, ref_fx):
    for key, th in _mvn_shapes():
        np.testing.assert_allclose(oracle.setup_mvn_sampler(th), ref_fx[key], rtol=1e-12, atol=1e-15)
    c = _case()
    o = oracle.sample_predictive_priors(31, 40000, c["w"], c["theta"], c["ptype"], c["pa"], c["pb"], c["dv"])
    _check_against_reference_samples(o["samples"], ref_fx["indep_samples"], c["theta"])
    rs = ref_fx["indep_samples"]
    assert rs[:, 0].min() >= 0.0 and rs[:, 0].max() <= 1.0 and np.all(rs[:, 1] == np.round(rs[:, 1]))   # the reference's validity + recast
    m = _mvn_case()
    np.testing.assert_allclose(oracle.setup_mvn_sampler(m["theta"]), ref_fx["mvn_L"], rtol=1e-12, atol=1e-15)
    o = oracle.sample_mvn_predictive_priors(32, 40000, m["w"], m["theta"], m["ptype"], m["pa"], m["pb"], ref_fx["mvn_L"])
    _check_against_reference_samples(o["samples"], ref_fx["mvn_samples"], m["theta"])
    assert np.all(ref_fx["indep_fallback"] == 0.5)                               # Priors.h:26-28 as the reference executes it


@pytest.mark.gpu
def test_cuda_sampling_matches_reference_fixture(api, ref_fx):
    for key, th in _mvn_shapes():
        np.testing.assert_allclose(api.setup_mvn_sampler(th), ref_fx[key], rtol=1e-10, atol=1e-13)
    c = _case()
    g = api.sample_predictive_priors(515, 40000, c["w"], c["theta"], c["dv"], c["lo"], c["hi"], c["mean"], integral=c["integral"])
    _check_against_reference_samples(g, ref_fx["indep_samples"], c["theta"])
    m = _mvn_case()
    g = api.sample_mvn_predictive_priors(516, 40000, m["w"], m["theta"], ref_fx["mvn_L"], m["lo"], m["hi"])
    _check_against_reference_samples(g, ref_fx["mvn_samples"], m["theta"])
