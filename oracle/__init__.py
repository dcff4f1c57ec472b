"""ctypes loader for the CPU oracle (oracle/abc_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. The product package (abcsmc_b200) never imports this.
All matrices are numpy float64, Fortran (column-major) order, matching Eigen::MatrixXd.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

_dp = C.POINTER(C.c_double)
_up = C.POINTER(C.c_uint64)


def build(force=False):
    """Compile oracle/liboracle.so with the reference's flags (see oracle/Makefile)."""
    src = os.path.join(_HERE, "abc_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.orc_normalcdf.restype = C.c_double
        L.orc_normalcdf.argtypes = [C.c_double]
        L.orc_wilcoxon.restype = C.c_double
        L.orc_gsl_ran_gaussian_pdf.restype = C.c_double
        L.orc_gsl_ran_gaussian_pdf.argtypes = [C.c_double, C.c_double]
        L.orc_prior_likelihood.restype = C.c_double
        L.orc_prior_likelihood.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
        L.orc_pls_fit.restype = C.c_void_p
        L.orc_pls_cv_new_data.restype = C.c_void_p
        L.orc_pls_cv_loo.restype = C.c_void_p
        L.orc_pls_cv_lso.restype = C.c_void_p
        L.orc_residual_rows.restype = C.c_long
        _lib = L
    return _lib


def _f(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _p(a):
    return a.ctypes.data_as(_dp)


def _u(a):
    return a.ctypes.data_as(_up)


def colwise_mean(X):
    X = _f(X); out = np.empty(X.shape[1])
    lib().orc_colwise_mean(_p(X), C.c_long(X.shape[0]), C.c_long(X.shape[1]), _p(out)); return out


def colwise_stdev(X, mean=None):
    X = _f(X); mean = colwise_mean(X) if mean is None else _f(mean); out = np.empty(X.shape[1])
    lib().orc_colwise_stdev(_p(X), C.c_long(X.shape[0]), C.c_long(X.shape[1]), _p(mean), _p(out)); return out


def colwise_z_scores(X, mean=None, sd=None):
    X = _f(X); Z = np.empty_like(X, order="F")
    if mean is None:
        lib().orc_colwise_z_scores_auto(_p(X), C.c_long(X.shape[0]), C.c_long(X.shape[1]), _p(Z))
    else:
        mean = _f(mean); sd = _f(sd)
        lib().orc_colwise_z_scores(_p(X), C.c_long(X.shape[0]), C.c_long(X.shape[1]), _p(mean), _p(sd), _p(Z))
    return Z


def z_scores(obs, mean, sd):
    obs, mean, sd = _f(obs), _f(mean), _f(sd); out = np.empty_like(obs)
    lib().orc_z_scores(_p(obs), _p(mean), _p(sd), C.c_long(obs.size), _p(out)); return out


def normalcdf(z):
    return lib().orc_normalcdf(float(z))


def wilcoxon(e1, e2):
    e1, e2 = _f(e1), _f(e2)
    return lib().orc_wilcoxon(_p(e1), _p(e2), C.c_long(e1.size))


def ordered(v):
    v = _f(v); out = np.empty(v.size, dtype=np.uint64)
    lib().orc_ordered(_p(v), C.c_long(v.size), _u(out)); return out


def euclidean(S, ref):
    S, ref = _f(S), _f(ref); out = np.empty(S.shape[0])
    lib().orc_euclidean(_p(S), C.c_long(S.shape[0]), C.c_long(S.shape[1]), _p(ref), _p(out)); return out


def dominant_eigenvector_sym(S):
    S = _f(S); out = np.empty(S.shape[0])
    lib().orc_dominant_eigenvector_sym(_p(S), C.c_long(S.shape[0]), _p(out)); return out


KERNEL_TYPE1, KERNEL_TYPE2 = 0, 1
RESS, MSE = 0, 1


class Residual:
    """PLS::Residual (lib/PLS/include/PLS/pls.h:44-53): M error matrices n_obs x A."""

    def __init__(self, handle, M, A):
        self._h, self.M, self.A = C.c_void_p(handle), M, A
        self.n = lib().orc_residual_rows(self._h)

    def errors(self):
        out = np.empty((self.M, self.A, self.n))
        lib().orc_residual_errors(self._h, _p(out))
        return [np.asfortranarray(out[y].T) for y in range(self.M)]

    def validation(self, out_type=RESS):
        out = np.empty((self.M, self.A), order="F")
        lib().orc_validation(self._h, C.c_int(out_type), _p(out)); return out

    def optimal_num_components(self, alpha=0.1):
        out = np.empty(self.M, dtype=np.uint64)
        lib().orc_optimal_num_components(self._h, C.c_double(alpha), _u(out)); return out

    def __del__(self):
        if _lib is not None and self._h:
            _lib.orc_residual_free(self._h); self._h = None


class Model:
    """PLS::Model (lib/PLS/include/PLS/pls.h:184-266)."""

    def __init__(self, X, Y, method=KERNEL_TYPE1, max_components=None):
        X, Y = _f(X), _f(Y)
        if Y.ndim == 1:
            Y = _f(Y.reshape(-1, 1))
        self.N, self.K = X.shape; self.M = Y.shape[1]
        self.A = self.K if max_components is None else int(max_components)
        self.method = method
        self._h = C.c_void_p(lib().orc_pls_fit(_p(X), _p(Y), C.c_long(self.N), C.c_long(self.K), C.c_long(self.M),
                                               C.c_int(method), C.c_long(self.A)))

    def _get(self, which, rows):
        out = np.empty((rows, self.A), order="F")
        lib().orc_pls_get(self._h, C.c_char(which.encode()), _p(out)); return out

    @property
    def P(self): return self._get("P", self.K)
    @property
    def W(self): return self._get("W", self.K)
    @property
    def R(self): return self._get("R", self.K)
    @property
    def Q(self): return self._get("Q", self.M)
    @property
    def T(self): return self._get("T", self.N)

    def scores(self, Xn, comp=None):
        Xn = _f(np.atleast_2d(Xn)); comp = self.A if comp is None else int(comp)
        out = np.empty((Xn.shape[0], comp), order="F")
        lib().orc_pls_scores(self._h, _p(Xn), C.c_long(Xn.shape[0]), C.c_long(comp), _p(out)); return out

    def coefficients(self, comp=None):
        comp = self.A if comp is None else int(comp)
        out = np.empty((self.K, self.M), order="F")
        lib().orc_pls_coefficients(self._h, C.c_long(comp), _p(out)); return out

    def fitted_values(self, Xn, comp=None):
        Xn = _f(Xn); comp = self.A if comp is None else int(comp)
        out = np.empty((Xn.shape[0], self.M), order="F")
        lib().orc_pls_fitted_values(self._h, _p(Xn), C.c_long(Xn.shape[0]), C.c_long(comp), _p(out)); return out

    def residuals(self, Xn, Yn, comp=None):
        Xn, Yn = _f(Xn), _f(Yn); comp = self.A if comp is None else int(comp)
        out = np.empty((Xn.shape[0], self.M), order="F")
        lib().orc_pls_residuals(self._h, _p(Xn), _p(Yn), C.c_long(Xn.shape[0]), C.c_long(comp), _p(out)); return out

    def SSE(self, Xn, Yn, comp=None):
        Xn, Yn = _f(Xn), _f(Yn); comp = self.A if comp is None else int(comp)
        out = np.empty(self.M)
        lib().orc_pls_SSE(self._h, _p(Xn), _p(Yn), C.c_long(Xn.shape[0]), C.c_long(comp), _p(out)); return out

    def explained_variance(self, Xn, Yn, comp=None):
        Xn, Yn = _f(Xn), _f(Yn); comp = self.A if comp is None else int(comp)
        out = np.empty(self.M)
        lib().orc_pls_explained_variance(self._h, _p(Xn), _p(Yn), C.c_long(Xn.shape[0]), C.c_long(comp), _p(out)); return out

    def cv_NEW_DATA(self, Xn, Yn):
        Xn, Yn = _f(Xn), _f(Yn)
        return Residual(lib().orc_pls_cv_new_data(self._h, _p(Xn), _p(Yn), C.c_long(Xn.shape[0])), self.M, self.A)

    def cv_LOO(self):
        return Residual(lib().orc_pls_cv_loo(self._h), self.M, self.A)

    def cv_LSO(self, shuffles, test_size):
        """shuffles: (num_trials, N) row indices, each row the shuffled `full` vector of one trial (pls.cpp:218-227)"""
        sh = np.ascontiguousarray(shuffles, dtype=np.uint64)
        return Residual(lib().orc_pls_cv_lso(self._h, _u(sh), C.c_long(int(test_size)), C.c_long(sh.shape[0])), self.M, self.A)

    def __del__(self):
        if _lib is not None and self._h:
            _lib.orc_pls_free(self._h); self._h = None


def particle_ranking_PLS(met, par, target, training_fraction=0.5):
    """ABC::particle_ranking_PLS (src/AbcUtil.cpp:423-458). Returns dict with order (full N),
    dist, ncomp (per parameter), ncomp_used, press (P x K)."""
    met, par, target = _f(met), _f(par), _f(target)
    N, K = met.shape; P = par.shape[1]
    order = np.empty(N, dtype=np.uint64); dist = np.empty(N); ncomp = np.empty(P, dtype=np.uint64)
    used = C.c_uint64(0); press = np.empty((P, K), order="F")
    lib().orc_particle_ranking_PLS(_p(met), _p(par), C.c_long(N), C.c_long(K), C.c_long(P), _p(target),
                                   C.c_double(training_fraction), _u(order), _p(dist), _u(ncomp), C.byref(used), _p(press))
    return dict(order=order, dist=dist, ncomp=ncomp, ncomp_used=int(used.value), press=press)


def particle_ranking_simple(met, target):
    met, target = _f(met), _f(target)
    N, K = met.shape
    order = np.empty(N, dtype=np.uint64); dist = np.empty(N)
    lib().orc_particle_ranking_simple(_p(met), C.c_long(N), C.c_long(K), _p(target), _u(order), _p(dist))
    return dict(order=order, dist=dist)


def calculate_doubled_variance(params):
    params = _f(params); out = np.empty(params.shape[1])
    lib().orc_calculate_doubled_variance(_p(params), C.c_long(params.shape[0]), C.c_long(params.shape[1]), _p(out)); return out


def gsl_ran_gaussian_pdf(x, sigma):
    return lib().orc_gsl_ran_gaussian_pdf(float(x), float(sigma))


PRIOR_UNIFORM, PRIOR_DISCRETE_UNIFORM, PRIOR_GAUSSIAN = 0, 1, 2


def prior_likelihood(ptype, a, b, v):
    return lib().orc_prior_likelihood(int(ptype), float(a), float(b), float(v))


def weight_predictive_prior0(n):
    out = np.empty(n); lib().orc_weight_predictive_prior0(C.c_long(n), _p(out)); return out


def weight_predictive_prior(numer, params, prev_params, prev_w, prev_dv):
    numer, params, prev_params, prev_w, prev_dv = map(_f, (numer, params, prev_params, prev_w, prev_dv))
    out = np.empty(params.shape[0])
    lib().orc_weight_predictive_prior(_p(numer), _p(params), C.c_long(params.shape[0]), _p(prev_params),
                                      C.c_long(prev_params.shape[0]), _p(prev_w), _p(prev_dv), C.c_long(params.shape[1]), _p(out))
    return out


def sample_predictive_priors(seed, num_samples, weights, parameter_prior, ptype, pa, pb, doubled_variance, max_attempts=1000):
    """ABC::sample_predictive_priors restated (src/AbcUtil.cpp:378-390) on a splitmix64 stream: the distributional checker.
    ptype/pa/pb: prior type and its two numbers per parameter (as prior_likelihood). Returns samples, parent rows, fall-backs."""
    w, th, pa, pb, dv = map(_f, (weights, parameter_prior, pa, pb, doubled_variance))
    pt = np.ascontiguousarray(np.asarray(ptype, dtype=np.int32))
    n_pp, P = th.shape
    out = np.empty((int(num_samples), P), order="F")
    parent = np.empty(int(num_samples), dtype=np.uint64)
    fb = C.c_long(0)
    lib().orc_sample_predictive_priors(C.c_uint64(int(seed)), C.c_long(int(num_samples)), _p(w), _p(th), C.c_long(n_pp), C.c_long(P),
                                       pt.ctypes.data_as(C.c_void_p), _p(pa), _p(pb), _p(dv), C.c_long(int(max_attempts)), _p(out),
                                       parent.ctypes.data_as(C.c_void_p), C.byref(fb))
    return {"samples": out, "parent": parent, "fallbacks": int(fb.value)}


def setup_mvn_sampler(params):
    """ABC::setup_mvn_sampler restated (src/AbcUtil.cpp:462-488): lower Cholesky factor, P x P."""
    th = _f(params)
    n_pp, P = th.shape
    L = np.empty((P, P), order="F")
    lib().orc_setup_mvn_sampler.restype = C.c_int
    if lib().orc_setup_mvn_sampler(_p(th), C.c_long(n_pp), C.c_long(P), _p(L)) != 0:
        raise ValueError("covariance not positive definite")
    return L


def sample_mvn_predictive_priors(seed, num_samples, weights, parameter_prior, ptype, pa, pb, L, max_attempts=100000):
    """ABC::sample_mvn_predictive_priors restated (src/AbcUtil.cpp:392-404) on a splitmix64 stream: the distributional checker."""
    w, th, pa, pb, L = map(_f, (weights, parameter_prior, pa, pb, L))
    pt = np.ascontiguousarray(np.asarray(ptype, dtype=np.int32))
    n_pp, P = th.shape
    out = np.empty((int(num_samples), P), order="F")
    parent = np.empty(int(num_samples), dtype=np.uint64)
    fl = C.c_long(0)
    lib().orc_sample_mvn_predictive_priors(C.c_uint64(int(seed)), C.c_long(int(num_samples)), _p(w), _p(th), C.c_long(n_pp), C.c_long(P),
                                           pt.ctypes.data_as(C.c_void_p), _p(pa), _p(pb), _p(L), C.c_long(int(max_attempts)), _p(out),
                                           parent.ctypes.data_as(C.c_void_p), C.byref(fl))
    return {"samples": out, "parent": parent, "failures": int(fl.value)}
