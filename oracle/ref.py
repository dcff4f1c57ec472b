"""ctypes binding of oracle/_ref/libabcref.so: the reference's OWN sources (lib/PLS/src/pls.cpp, src/AbcUtil.cpp), unmodified,
compiled against the Eigen / GSL stand-ins of oracle/shim/ (see oracle/Makefile, oracle/ref_harness.cpp).

TEST INFRASTRUCTURE ONLY, and only where /root/reference exists (the authoring container): tests/test_ref_pin.py uses it to pin
oracle/abc_oracle.cpp, tests/golden/make_ref_fixtures.py to write the fixtures that travel. Nothing in the product, the GPU
tests, smoke() or bench.py imports this module.

The module re-executes oracle/__init__.py (the oracle's ctypes wrappers) against this library: every orc_<name> call is served by
ref_<name>, which has the same signature (ref_harness.cpp). Entry points whose reference signature differs are defined here.
"""
import ctypes as C
import importlib.util
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libabcref.so")
REFERENCE_ROOT = os.environ.get("ABC_REFERENCE_ROOT", "/root/reference")


def available():
    """True when the library is built or can be built (the reference's sources are present)."""
    return os.path.exists(LIB_PATH) or os.path.exists(os.path.join(REFERENCE_ROOT, "lib", "PLS", "src", "pls.cpp"))


def build(force=False):
    if not os.path.exists(os.path.join(REFERENCE_ROOT, "lib", "PLS", "src", "pls.cpp")):
        if os.path.exists(LIB_PATH):
            return LIB_PATH
        raise RuntimeError("reference sources not found under %s" % REFERENCE_ROOT)
    subprocess.check_call(["make", "-C", _HERE, "ref", "REF=" + REFERENCE_ROOT] + (["-B"] if force else ["-s"]), stdout=subprocess.DEVNULL)
    return LIB_PATH


class _Renamed:
    """Serves orc_<name> from ref_<name>."""

    def __init__(self, cdll):
        self._c = cdll

    def __getattr__(self, name):
        if not name.startswith("orc_"):
            raise AttributeError(name)
        return getattr(self._c, "ref_" + name[4:])


_cdll = None
_binding = None


def _load():
    global _cdll, _binding
    if _binding is None:
        build()
        _cdll = C.CDLL(LIB_PATH)
        for fn in ("ref_normalcdf", "ref_wilcoxon", "ref_prior_likelihood", "ref_calculate_nrmse", "ref_median"):
            getattr(_cdll, fn).restype = C.c_double
        _cdll.ref_normalcdf.argtypes = [C.c_double]
        _cdll.ref_prior_likelihood.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
        for fn in ("ref_pls_fit", "ref_pls_cv_new_data", "ref_pls_cv_loo", "ref_pls_cv_lso_seeded"):
            getattr(_cdll, fn).restype = C.c_void_p
        _cdll.ref_residual_rows.restype = C.c_long
        spec = importlib.util.spec_from_file_location("oracle._bound_to_reference", os.path.join(_HERE, "__init__.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        proxy = _Renamed(_cdll)
        mod._lib = proxy
        mod.lib = lambda: proxy
        _binding = mod
    return _binding


_dp = C.POINTER(C.c_double)


def _f(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _p(a):
    return a.ctypes.data_as(_dp)


def __getattr__(name):
    # colwise_mean, colwise_stdev, colwise_z_scores, z_scores, normalcdf, wilcoxon, ordered, euclidean, Model, Residual,
    # calculate_doubled_variance, prior_likelihood, KERNEL_TYPE1/2, RESS/MSE, PRIOR_*: the oracle's wrappers, bound to the reference
    if name.startswith("__"):
        raise AttributeError(name)
    return getattr(_load(), name)


def particle_ranking_PLS(met, par, target, training_fraction=0.5):
    """ABC::particle_ranking_PLS (src/AbcUtil.cpp:423-458): the full order is all the reference returns."""
    _load()
    met, par, target = _f(met), _f(par), _f(target)
    N, K = met.shape
    order = np.empty(N, dtype=np.uint64)
    _cdll.ref_particle_ranking_PLS(_p(met), _p(par), C.c_long(N), C.c_long(K), C.c_long(par.shape[1]), _p(target),
                                   C.c_double(training_fraction), order.ctypes.data_as(C.c_void_p))
    return order


def particle_ranking_simple(met, target):
    _load()
    met, target = _f(met), _f(target)
    N, K = met.shape
    order = np.empty(N, dtype=np.uint64)
    _cdll.ref_particle_ranking_simple(_p(met), C.c_long(N), C.c_long(K), _p(target), order.ctypes.data_as(C.c_void_p))
    return order


def weight_predictive_prior0(n, P=1):
    _load()
    out = np.empty(n)
    _cdll.ref_weight_predictive_prior0(C.c_long(n), C.c_long(P), _p(out))
    return out


def weight_predictive_prior(ptype, pa, pb, params, prev_params, prev_w, prev_dv):
    """ABC::weight_predictive_prior (src/AbcUtil.cpp:547-586) with Parameter objects built from (kind, a, b) per parameter:
    0 ContinuousUniformPrior(a, b), 1 DiscreteUniformPrior(a, b), 2 GaussianPrior(mean a, sd b) (include/AbcSmc/Priors.h)."""
    _load()
    pa, pb, params, prev_params, prev_w, prev_dv = map(_f, (pa, pb, params, prev_params, prev_w, prev_dv))
    pt = np.ascontiguousarray(np.asarray(ptype, dtype=np.int32))
    out = np.empty(params.shape[0])
    _cdll.ref_weight_predictive_prior(pt.ctypes.data_as(C.c_void_p), _p(pa), _p(pb), _p(params), C.c_long(params.shape[0]), _p(prev_params),
                                      C.c_long(prev_params.shape[0]), _p(prev_w), _p(prev_dv), C.c_long(params.shape[1]), _p(out))
    return out


def cv_LSO_seeded(model, test_fraction, num_trials, seed):
    """Model::cv_LSO (lib/PLS/src/pls.cpp:512-549) with std::mt19937(seed). `model` is a Model of this module."""
    b = _load()
    h = _cdll.ref_pls_cv_lso_seeded(model._h, C.c_double(test_fraction), C.c_long(int(num_trials)), C.c_uint32(int(seed)))
    return b.Residual(h, model.M, model.A)


def lso_shuffles(seed, N, test_size, num_trials):
    """The shuffled `full` vector of every trial of cv_LSO_seeded with the same seed (PLS::rand_nchoosek, pls.cpp:218-227)."""
    _load()
    out = np.empty((int(num_trials), int(N)), dtype=np.uint64)
    _cdll.ref_lso_shuffles(C.c_uint32(int(seed)), C.c_long(int(N)), C.c_long(int(test_size)), C.c_long(int(num_trials)), out.ctypes.data_as(C.c_void_p))
    return out


def calculate_nrmse(mets, observed):
    _load()
    mets, observed = _f(mets), _f(observed)
    return _cdll.ref_calculate_nrmse(_p(mets), C.c_long(mets.shape[0]), C.c_long(mets.shape[1]), _p(observed))


def median(v):
    _load()
    v = _f(v)
    return _cdll.ref_median(_p(v), C.c_long(v.size))


def setup_mvn_sampler(params):
    """ABC::setup_mvn_sampler (src/AbcUtil.cpp:462-488) on the stand-in's vcov + Cholesky (deterministic): lower factor, P x P."""
    _load()
    th = _f(params)
    n_pp, P = th.shape
    L = np.zeros((P, P), order="F")
    _cdll.ref_setup_mvn_sampler(_p(th), C.c_long(n_pp), C.c_long(P), _p(L))
    return L


def _priors(ptype, pa, pb):
    return np.ascontiguousarray(np.asarray(ptype, dtype=np.int32)), _f(pa), _f(pb)


def sample_predictive_priors(seed, num_samples, weights, parameter_prior, ptype, pa, pb, doubled_variance):
    """ABC::sample_predictive_priors (src/AbcUtil.cpp:378-390, Priors.h:18-41) on the stand-in's MT19937 stream."""
    _load()
    w, th, dv = _f(weights), _f(parameter_prior), _f(doubled_variance)
    pt, pa, pb = _priors(ptype, pa, pb)
    out = np.empty((int(num_samples), th.shape[1]), order="F")
    _cdll.ref_sample_predictive_priors(C.c_uint32(int(seed)), C.c_long(int(num_samples)), _p(w), _p(th), C.c_long(th.shape[0]), C.c_long(th.shape[1]),
                                       pt.ctypes.data_as(C.c_void_p), _p(pa), _p(pb), _p(dv), _p(out))
    return out


def sample_mvn_predictive_priors(seed, num_samples, weights, parameter_prior, ptype, pa, pb, L):
    """ABC::sample_mvn_predictive_priors (src/AbcUtil.cpp:392-404) on the stand-in's MT19937 stream; L as setup_mvn_sampler returns it."""
    _load()
    w, th, L = _f(weights), _f(parameter_prior), _f(L)
    pt, pa, pb = _priors(ptype, pa, pb)
    out = np.empty((int(num_samples), th.shape[1]), order="F")
    _cdll.ref_sample_mvn_predictive_priors(C.c_uint32(int(seed)), C.c_long(int(num_samples)), _p(w), _p(th), C.c_long(th.shape[0]), C.c_long(th.shape[1]),
                                           pt.ctypes.data_as(C.c_void_p), _p(pa), _p(pb), _p(L), _p(out))
    return out
