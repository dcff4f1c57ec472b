"""Host-side mirror of the reference's interface for the hot path, on top of the C ABI.

Same names and argument meaning as the reference:
  ABC::particle_ranking_PLS / particle_ranking_simple   src/AbcUtil.cpp:408-458
  ABC::calculate_doubled_variance                        src/AbcUtil.cpp:528-537
  ABC::weight_predictive_prior (both overloads)          src/AbcUtil.cpp:539-586
  ABC::euclidean                                         src/AbcUtil.cpp:320-324
  PLS::ordered / colwise_z_scores / colwise_stdev / wilcoxon / Model   lib/PLS/include/PLS/pls.h
Matrices are numpy float64; they are converted to column-major (Eigen::MatrixXd layout) if needed.
All arithmetic runs in the CUDA library; nothing here computes on the CPU and nothing falls back.
"""
import ctypes as C

import numpy as np

from . import _capi

KERNEL_TYPE1, KERNEL_TYPE2, KERNEL_TYPE1_STREAM = 0, 1, 2
RESS, MSE = 0, 1
KERNELS = ("pls_gram_kernel", "gram_kernel", "screen1_kernel", "screen2_kernel", "press_chk_kernel", "xb_kernel<0>", "xb_kernel<1>",
           "weights_main_kernel", "zscore_kernel", "pls_loo_kernel")
STAGES = ("moments_zscore", "pls_fit", "holdout_press", "wilcoxon_select", "project_distance", "ordering",
          "doubled_variance", "weight_update", "h2d", "d2h")


class Context:
    """One CUDA context wrapper per device (abcb200_ctx). Not thread-safe."""

    def __init__(self, device=0):
        self._lib = _capi.lib()
        h = C.c_void_p()
        rc = self._lib.abcb200_create(int(device), C.byref(h))
        if rc != 0:
            raise _capi.Abcb200Error(rc, "abcb200_create failed (no usable sm_100 CUDA device?) - there is no CPU fallback")
        self._h = h
        self.device = int(device)
        self._stream = None

    def check(self, rc):
        if rc != 0:
            raise _capi.Abcb200Error(rc, self._lib.abcb200_last_error(self._h).decode())

    def set_stream(self, cuda_stream):
        if cuda_stream != self._stream:
            self.check(self._lib.abcb200_set_stream(self._h, C.c_void_p(cuda_stream)))
            self._stream = cuda_stream

    def synchronize(self):
        self.check(self._lib.abcb200_synchronize(self._h))

    @property
    def launches(self):
        return int(self._lib.abcb200_launch_count(self._h))

    @property
    def exact_tests(self):
        return int(self._lib.abcb200_exact_test_count(self._h))

    def stat(self, which):
        """0 launches, 1 signed-rank tests of the last selection, 2 tests that reached level 2, 3 exact tests so far,
        4 component loop of the last PLS fit (1 pls_defl_kernel, 2 pls_gram_kernel, 3 pls_wide.cu)"""
        return int(self._lib.abcb200_stat(self._h, int(which)))

    def set_timers(self, stages=False, kernels=()):
        """CUDA-event instrumentation (off by default: every record costs launch path). `kernels`: names from KERNELS, or "all"."""
        mask = 0xFFFFFFFF if kernels == "all" else sum(1 << KERNELS.index(k) for k in kernels)
        self.check(self._lib.abcb200_set_timers(self._h, int(bool(stages)), mask))

    def set_tie_order(self, mode):
        """Placement of exact distance ties: TIES_BY_INDEX (default, ascending particle index) or TIES_STDSORT (where libstdc++'s
        std::sort leaves them in PLS::ordered, lib/PLS/include/PLS/pls.h:58-69; a host pass only when ties reach the output)."""
        self.check(self._lib.abcb200_set_tie_order(self._h, int(mode)))

    @property
    def tie_resorts(self):
        return int(self._lib.abcb200_stat(self._h, 8))

    def stage_ms(self):
        return {name: float(self._lib.abcb200_stage_ms(self._h, i)) for i, name in enumerate(STAGES)}

    def kernel_ms(self):
        return {name: float(self._lib.abcb200_kernel_ms(self._h, i)) for i, name in enumerate(KERNELS)}

    def close(self):
        if self._h:
            self._lib.abcb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


TIES_BY_INDEX, TIES_STDSORT = 0, 1


def tie_order_stdsort(dist, order):
    """abcb200_tie_order_stdsort: the host step of TIES_STDSORT on its own (no GPU). Returns (order, changed)."""
    dist = np.ascontiguousarray(np.asarray(dist, dtype=np.float64))
    order = np.ascontiguousarray(np.asarray(order, dtype=np.uint64)).copy()
    rc = _capi.lib().abcb200_tie_order_stdsort(dist.ctypes.data_as(C.c_void_p), dist.size, order.size, order.ctypes.data_as(C.c_void_p))
    if rc < 0:
        raise _capi.Abcb200Error(rc, "abcb200_tie_order_stdsort: bad argument")
    return order, bool(rc)


_default = {}


def get_context(device=0):
    if device not in _default:
        _default[device] = Context(device)
    return _default[device]


def _f(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _vec(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel())


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def particle_ranking_PLS(X_orig, Y_orig, target_values, training_fraction, top_n=0, method=KERNEL_TYPE1, ctx=None,
                         return_info=False):
    """ABC::particle_ranking_PLS(metrics, params, target, training_fraction) -> order (uint64).
    top_n > 0 returns only the leading top_n entries (what AbcSmc.cpp:645-646 keeps)."""
    ctx = ctx or get_context()
    met, par, tgt = _f(X_orig), _f(Y_orig), _vec(target_values)
    N, K = met.shape
    P = par.shape[1]
    if par.shape[0] != N or tgt.size != K:
        raise ValueError("shape mismatch")
    n_out = N if top_n <= 0 or top_n > N else int(top_n)
    order = np.empty(n_out, dtype=np.uint64)
    dist = np.empty(N) if return_info else None
    used = C.c_int(0)
    ncomp = np.zeros(P, dtype=np.int32)
    ctx.check(ctx._lib.abcb200_rank_pls(ctx._h, _ptr(met), N, _ptr(par), N, N, K, P, _ptr(tgt), float(training_fraction),
                                        int(method), n_out, _ptr(order), _ptr(dist), C.cast(C.byref(used), C.c_void_p), _ptr(ncomp)))
    if return_info:
        return dict(order=order, dist=dist, ncomp=ncomp, ncomp_used=used.value)
    return order


def particle_ranking_simple(X_orig, Y_orig, target_values, top_n=0, ctx=None, return_info=False):
    """ABC::particle_ranking_simple(metrics, params (unused), target) -> order."""
    ctx = ctx or get_context()
    met, tgt = _f(X_orig), _vec(target_values)
    N, K = met.shape
    n_out = N if top_n <= 0 or top_n > N else int(top_n)
    order = np.empty(n_out, dtype=np.uint64)
    dist = np.empty(N) if return_info else None
    ctx.check(ctx._lib.abcb200_rank_simple(ctx._h, _ptr(met), N, N, K, _ptr(tgt), n_out, _ptr(order), _ptr(dist)))
    if return_info:
        return dict(order=order, dist=dist)
    return order


def calculate_doubled_variance(params, ctx=None):
    """ABC::calculate_doubled_variance(params) -> Row of 2 * sample variance per column."""
    ctx = ctx or get_context()
    p = _f(params)
    out = np.empty(p.shape[1])
    ctx.check(ctx._lib.abcb200_doubled_variance(ctx._h, _ptr(p), p.shape[0], p.shape[0], p.shape[1], _ptr(out)))
    return out


def sample_predictive_priors(seed, num_samples, weights, parameter_prior, doubled_variance, lo, hi, prior_mean, integral=None,
                             max_attempts=1000, return_info=False, ctx=None):
    """ABC::sample_predictive_priors(RNG, num_samples, weights, parameter_prior, pars, doubled_variance) (src/AbcUtil.cpp:378-390)
    with the Parameter objects flattened to (lo, hi, integral, prior_mean) and the gsl_rng replaced by a 64-bit seed
    (distributional parity, include/abcsmc_b200.h). Returns the num_samples x P proposals; with return_info also the
    parent row of every sample and the number of prior-mean fall-backs."""
    ctx = ctx or get_context()
    th = _f(parameter_prior)
    n_pp, P = th.shape
    w, dv, lo, hi, mean = _vec(weights), _vec(doubled_variance), _vec(lo), _vec(hi), _vec(prior_mean)
    if w.size != n_pp or dv.size != P or lo.size != P or hi.size != P or mean.size != P:
        raise ValueError("shape mismatch")
    integ = None if integral is None else np.ascontiguousarray(np.asarray(integral, dtype=np.int32))
    out = np.empty((int(num_samples), P), order="F")
    parent = np.empty(int(num_samples), dtype=np.uint64)
    fb = np.zeros(1, dtype=np.uint64)
    ctx.check(ctx._lib.abcb200_sample_predictive_priors(ctx._h, int(seed) & 0xFFFFFFFFFFFFFFFF, int(num_samples), _ptr(w), _ptr(th), n_pp, n_pp, P,
                                                        _ptr(dv), _ptr(lo), _ptr(hi), _ptr(integ), _ptr(mean), int(max_attempts), _ptr(out),
                                                        int(num_samples), _ptr(parent), _ptr(fb)))
    if return_info:
        return {"samples": out, "parent": parent, "fallbacks": int(fb[0])}
    return out


def setup_mvn_sampler(params, ctx=None):
    """ABC::setup_mvn_sampler(params) (src/AbcUtil.cpp:462-488): lower Cholesky factor (P x P) of the sample covariance with a doubled diagonal."""
    ctx = ctx or get_context()
    th = _f(params)
    n_pp, P = th.shape
    L = np.empty((P, P), order="F")
    ctx.check(ctx._lib.abcb200_setup_mvn_sampler(ctx._h, _ptr(th), n_pp, n_pp, P, _ptr(L)))
    return L


def sample_mvn_predictive_priors(seed, num_samples, weights, parameter_prior, L, lo, hi, integral=None, max_attempts=100000,
                                 return_info=False, ctx=None):
    """ABC::sample_mvn_predictive_priors(RNG, num_samples, weights, parameter_prior, pars, L) (src/AbcUtil.cpp:392-404); L from
    setup_mvn_sampler. Distributional parity (include/abcsmc_b200.h)."""
    ctx = ctx or get_context()
    th = _f(parameter_prior)
    n_pp, P = th.shape
    w, lo, hi, Lf = _vec(weights), _vec(lo), _vec(hi), _f(L)
    if w.size != n_pp or lo.size != P or hi.size != P or Lf.shape != (P, P):
        raise ValueError("shape mismatch")
    integ = None if integral is None else np.ascontiguousarray(np.asarray(integral, dtype=np.int32))
    out = np.empty((int(num_samples), P), order="F")
    parent = np.empty(int(num_samples), dtype=np.uint64)
    fl = np.zeros(1, dtype=np.uint64)
    ctx.check(ctx._lib.abcb200_sample_mvn_predictive_priors(ctx._h, int(seed) & 0xFFFFFFFFFFFFFFFF, int(num_samples), _ptr(w), _ptr(th), n_pp, n_pp, P,
                                                            _ptr(Lf), _ptr(lo), _ptr(hi), _ptr(integ), int(max_attempts), _ptr(out), int(num_samples),
                                                            _ptr(parent), _ptr(fl)))
    if return_info:
        return {"samples": out, "parent": parent, "failures": int(fl[0])}
    return out


def weight_predictive_prior(numer, params, prev_params=None, prev_weights=None, prev_doubled_variance=None, algo=0, ctx=None):
    """ABC::weight_predictive_prior. With only `params`: set 0, uniform 1/N (numer ignored).
    Otherwise numer[i] = prod_p prior_p.likelihood(params[i,p]) (None = all ones), as computed by the caller's
    Parameter objects (src/AbcUtil.cpp:559-561)."""
    ctx = ctx or get_context()
    th = _f(params)
    n_new, P = th.shape
    out = np.empty(n_new)
    if prev_params is None:
        ctx.check(ctx._lib.abcb200_weights_set0(ctx._h, n_new, _ptr(out)))
        return out
    tho, wo, dv = _f(prev_params), _vec(prev_weights), _vec(prev_doubled_variance)
    nm = None if numer is None else _vec(numer)
    if tho.shape[1] != P or wo.size != tho.shape[0] or dv.size != P or (nm is not None and nm.size != n_new):
        raise ValueError("shape mismatch")
    ctx.check(ctx._lib.abcb200_weights(ctx._h, _ptr(nm), _ptr(th), n_new, n_new, _ptr(tho), tho.shape[0], tho.shape[0], _ptr(wo),
                                       _ptr(dv), P, int(algo), _ptr(out)))
    return out


def colwise_moments(X, ctx=None):
    ctx = ctx or get_context()
    x = _f(X)
    mean, sd = np.empty(x.shape[1]), np.empty(x.shape[1])
    ctx.check(ctx._lib.abcb200_colwise_moments(ctx._h, _ptr(x), x.shape[0], x.shape[0], x.shape[1], _ptr(mean), _ptr(sd)))
    return mean, sd


def colwise_stdev(X, ctx=None):
    """PLS::colwise_stdev(mat)"""
    return colwise_moments(X, ctx)[1]


def colwise_z_scores(X, mean=None, stdev=None, ctx=None):
    """PLS::colwise_z_scores(mat[, mean, stdev])"""
    ctx = ctx or get_context()
    x = _f(X)
    z = np.empty_like(x, order="F")
    m = None if mean is None else _vec(mean)
    s = None if stdev is None else _vec(stdev)
    if (m is not None and m.size != x.shape[1]) or (s is not None and s.size != x.shape[1]):
        raise ValueError("mean / stdev must hold one value per column")
    ctx.check(ctx._lib.abcb200_colwise_z_scores(ctx._h, _ptr(x), x.shape[0], x.shape[0], x.shape[1], _ptr(m), _ptr(s), _ptr(z), x.shape[0]))
    return z


def gram(X, Y, ctx=None):
    """(X^T X, X^T Y): the products PLS::Model::plsr starts from (pls.cpp:396, :398)."""
    ctx = ctx or get_context()
    x, y = _f(X), _f(Y)
    K, M = x.shape[1], y.shape[1]
    if y.shape[0] != x.shape[0]:
        raise ValueError("X and Y must have the same number of rows")
    xx = np.empty((K, K), order="F"); xy = np.empty((K, M), order="F")
    ctx.check(ctx._lib.abcb200_gram(ctx._h, _ptr(x), x.shape[0], _ptr(y), y.shape[0], x.shape[0], K, M, _ptr(xx), _ptr(xy)))
    return xx, xy


def euclidean(sims, ref, ctx=None):
    """ABC::euclidean(sims, ref)"""
    ctx = ctx or get_context()
    s, r = _f(sims), _vec(ref)
    if r.size != s.shape[1]:
        raise ValueError("ref must hold one value per column of sims")
    out = np.empty(s.shape[0])
    ctx.check(ctx._lib.abcb200_euclidean(ctx._h, _ptr(s), s.shape[0], s.shape[0], s.shape[1], _ptr(r), _ptr(out)))
    return out


def ordered(v, top_n=0, ctx=None):
    """PLS::ordered(v): ascending index order (ties by ascending index, or std::sort's placement after ctx.set_tie_order(TIES_STDSORT));
    top_n > 0 returns only the leading entries."""
    ctx = ctx or get_context()
    x = _vec(v)
    n_out = x.size if top_n <= 0 or top_n > x.size else int(top_n)
    out = np.empty(n_out, dtype=np.uint64)
    ctx.check(ctx._lib.abcb200_ordered_top(ctx._h, _ptr(x), x.size, n_out, _ptr(out)))
    return out


def wilcoxon(err_1, err_2, ctx=None):
    """PLS::wilcoxon(err_1, err_2) -> p-value"""
    ctx = ctx or get_context()
    a, b = _vec(err_1), _vec(err_2)
    p = C.c_double(0)
    ctx.check(ctx._lib.abcb200_wilcoxon(ctx._h, _ptr(a), _ptr(b), a.size, C.cast(C.byref(p), C.c_void_p)))
    return p.value


class Residual:
    """PLS::Residual (lib/PLS/include/PLS/pls.h:44-53): one (rows x A) error matrix per response, held as the cube
    errors[y, c, i] the C ABI exchanges. validation / optimal_num_components are PLS::validation (pls.cpp:235-261) and
    PLS::optimal_num_components (:265-289) evaluated on the GPU."""

    def __init__(self, cube, method, ctx=None, press=None, ncomp=None, alpha=None):
        self.cube, self.method, self.ctx = np.ascontiguousarray(cube), method, ctx or get_context()
        self.M, self.A, self.n = self.cube.shape
        self._press, self._ncomp, self._alpha = press, ncomp, alpha

    def errors(self):
        return [np.asfortranarray(self.cube[y].T) for y in range(self.M)]

    def _select(self, out_type, alpha):
        press = np.empty((self.M, self.A), order="F")
        ncomp = np.zeros(self.M, dtype=np.int32)
        self.ctx.check(self.ctx._lib.abcb200_residual_select(self.ctx._h, _ptr(self.cube), self.n, self.M, self.A, int(out_type), float(alpha),
                                                             _ptr(press), _ptr(ncomp)))
        return press, ncomp

    def validation(self, out_type=RESS):
        if self._press is not None:
            return self._press / (self.n if out_type == MSE else 1.0)
        return self._select(out_type, 0.1)[0]

    def optimal_num_components(self, alpha=0.1):
        if self._ncomp is not None and alpha == self._alpha:
            return self._ncomp
        return self._select(RESS, alpha)[1]


class Model:
    """PLS::Model(X, Y, algorithm, max_components): fits on construction (lib/PLS/src/pls.cpp:340-359)."""

    def __init__(self, X, Y, algorithm=KERNEL_TYPE1, max_components=None, ctx=None):
        self.ctx = ctx or get_context()
        x, y = _f(X), _f(Y)
        if y.ndim == 1:
            y = _f(y.reshape(-1, 1))
        self.N, self.K = x.shape
        self.M = y.shape[1]
        self.A = self.K if max_components is None else int(max_components)
        self.method = int(algorithm)
        h = C.c_void_p()
        self.ctx.check(self.ctx._lib.abcb200_pls_fit(self.ctx._h, _ptr(x), self.N, _ptr(y), self.N, self.N, self.K, self.M,
                                                     self.method, self.A, C.byref(h)))
        self._h = h
        self._X, self._Y = x, y        # the reference's ctor copies them too (pls.cpp:344); cv_LOO / cv_LSO refit from them

    def _get(self, which, rows):
        out = np.empty((rows, self.A), order="F")
        self.ctx.check(self.ctx._lib.abcb200_pls_get(self._h, which.encode(), _ptr(out)))
        return out

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

    def scores(self, X_new, comp=None):
        x = _f(np.atleast_2d(X_new)); comp = self.A if comp is None else int(comp)
        out = np.empty((x.shape[0], comp), order="F")
        self.ctx.check(self.ctx._lib.abcb200_pls_scores(self._h, _ptr(x), x.shape[0], x.shape[0], comp, _ptr(out)))
        return out

    def coefficients(self, comp=None):
        comp = self.A if comp is None else int(comp)
        out = np.empty((self.K, self.M), order="F")
        self.ctx.check(self.ctx._lib.abcb200_pls_coefficients(self._h, comp, _ptr(out)))
        return out

    def fitted_values(self, X_new, comp=None):
        x = _f(X_new); comp = self.A if comp is None else int(comp)
        out = np.empty((x.shape[0], self.M), order="F")
        self.ctx.check(self.ctx._lib.abcb200_pls_fitted_values(self._h, _ptr(x), x.shape[0], x.shape[0], comp, _ptr(out)))
        return out

    def residuals(self, X_new, Y_new, comp=None):
        x, y = _f(X_new), _f(Y_new); comp = self.A if comp is None else int(comp)
        out = np.empty((x.shape[0], self.M), order="F")
        self.ctx.check(self.ctx._lib.abcb200_pls_residuals(self._h, _ptr(x), x.shape[0], _ptr(y), y.shape[0], x.shape[0], comp, _ptr(out)))
        return out

    def SSE(self, X_new, Y_new, comp=None):
        x, y = _f(X_new), _f(Y_new); comp = self.A if comp is None else int(comp)
        out = np.empty(self.M)
        self.ctx.check(self.ctx._lib.abcb200_pls_sse(self._h, _ptr(x), x.shape[0], _ptr(y), y.shape[0], x.shape[0], comp, _ptr(out)))
        return out

    def explained_variance(self, X_new, Y_new, comp=None):
        x, y = _f(X_new), _f(Y_new); comp = self.A if comp is None else int(comp)
        out = np.empty(self.M)
        self.ctx.check(self.ctx._lib.abcb200_pls_explained_variance(self._h, _ptr(x), x.shape[0], _ptr(y), y.shape[0], x.shape[0], comp, _ptr(out)))
        return out

    def cv_LOO(self, alpha=0.1):
        """Model::cv_LOO (pls.cpp:469-491): N refits as rank-one down-dates, batched over the SMs. Returns a Residual whose
        validation / optimal_num_components(alpha) were computed in the same call."""
        cube = np.empty((self.M, self.A, self.N))
        press = np.empty((self.M, self.A), order="F")
        ncomp = np.zeros(self.M, dtype=np.int32)
        self.ctx.check(self.ctx._lib.abcb200_pls_cv_loo(self.ctx._h, _ptr(self._X), self.N, _ptr(self._Y), self.N, self.N, self.K, self.M, self.A,
                                                        RESS, float(alpha), _ptr(cube), _ptr(press), _ptr(ncomp)))
        return Residual(cube, "LOO", self.ctx, press, ncomp, alpha)

    def cv_LSO(self, shuffles, test_size, alpha=0.1):
        """Model::cv_LSO (pls.cpp:512-549). shuffles: (num_trials, N) row indices, each row the shuffled `full` vector of one
        trial (rand_nchoosek, :218-227); the first N - test_size entries train, the rest are predicted."""
        sh = np.ascontiguousarray(shuffles, dtype=np.uint64)
        trials = sh.shape[0]
        if sh.shape[1] != self.N:
            raise ValueError("each shuffle must hold N indices")
        cube = np.empty((self.M, self.A, trials * int(test_size)))
        press = np.empty((self.M, self.A), order="F")
        ncomp = np.zeros(self.M, dtype=np.int32)
        self.ctx.check(self.ctx._lib.abcb200_pls_cv_lso(self.ctx._h, _ptr(self._X), self.N, _ptr(self._Y), self.N, self.N, self.K, self.M, self.A,
                                                        self.method, _ptr(sh), int(test_size), trials, RESS, float(alpha), _ptr(cube), _ptr(press),
                                                        _ptr(ncomp)))
        return Residual(cube, "LSO", self.ctx, press, ncomp, alpha)

    def cv_NEW_DATA(self, X_new, Y_new, out_type=RESS, alpha=0.1):
        """cv_NEW_DATA + validation + optimal_num_components, streamed: returns (press M x A, n_comp M)."""
        x, y = _f(X_new), _f(Y_new)
        press = np.empty((self.M, self.A), order="F")
        ncomp = np.zeros(self.M, dtype=np.int32)
        self.ctx.check(self.ctx._lib.abcb200_pls_cv_new_data(self._h, _ptr(x), x.shape[0], _ptr(y), y.shape[0], x.shape[0], int(out_type),
                                                             float(alpha), _ptr(press), _ptr(ncomp)))
        return press, ncomp

    def close(self):
        if getattr(self, "_h", None):
            self.ctx._lib.abcb200_pls_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


PRIOR_UNIFORM, PRIOR_DISCRETE_UNIFORM, PRIOR_GAUSSIAN = 0, 1, 2
FILTER_PLS, FILTER_SIMPLE = 0, 1


class SmcChain:
    """abcb200_chain: one call per SMC set, the previous set's predictive prior resident on the device.
    Mirrors what AbcSmc keeps per set (_predictive_prior, _doubled_variance, _weights: AbcSmc.cpp:634-664, 1041-1066)."""

    def __init__(self, n_params, ctx=None):
        self.ctx = ctx or get_context()
        self.P = int(n_params)
        h = C.c_void_p()
        self.ctx.check(self.ctx._lib.abcb200_chain_create(self.ctx._h, self.P, C.byref(h)))
        self._h = h

    @property
    def sets(self):
        return int(self.ctx._lib.abcb200_chain_sets(self._h))

    def process_set(self, metrics, params, target, top_n, filtering=FILTER_PLS, training_fraction=0.5, method=KERNEL_TYPE1,
                    priors=None, numer_all=None, report=True):
        """priors: (type, a, b) arrays of length P (PRIOR_*), or None. Returns dict(order, weights, doubled_variance, ncomp_used,
        nrmse, mean_par, mean_met, median_par, median_met)."""
        met, par, tgt = _f(metrics), _f(params), _vec(target)
        N, K = met.shape
        P = self.P
        if par.shape != (N, P) or tgt.size != K:
            raise ValueError("shape mismatch")
        n = N if top_n <= 0 or top_n > N else int(top_n)
        order = np.empty(n, dtype=np.uint64); w = np.empty(n); dv = np.empty(P)
        rep = np.empty(1 + 2 * (P + K)) if report else None
        used = C.c_int(0)
        pt = pa = pb = None
        if priors is not None:
            pt = np.ascontiguousarray(np.asarray(priors[0], dtype=np.int32)); pa = _vec(priors[1]); pb = _vec(priors[2])
        na = None if numer_all is None else _vec(numer_all)
        if (na is not None and na.size != N) or (pt is not None and not (pt.size == pa.size == pb.size == P)):
            raise ValueError("numer_all must hold N values, priors P values each")
        self.ctx.check(self.ctx._lib.abcb200_chain_process_set(self._h, _ptr(met), N, _ptr(par), N, N, K, _ptr(tgt), int(filtering), float(training_fraction),
                                                               int(method), n, _ptr(pt), _ptr(pa), _ptr(pb), _ptr(na), _ptr(order), _ptr(w), _ptr(dv), _ptr(rep),
                                                               C.cast(C.byref(used), C.c_void_p)))
        out = dict(order=order, weights=w, doubled_variance=dv, ncomp_used=used.value)
        if report:
            out.update(nrmse=rep[0], mean_par=rep[1:1 + P], mean_met=rep[1 + P:1 + P + K], median_par=rep[1 + P + K:1 + 2 * P + K],
                       median_met=rep[1 + 2 * P + K:])
        return out

    def state(self):
        """The last finished set as the device holds it: (theta n x P in rank order, weights n, doubled variance P)."""
        n = C.c_int64(0)
        self.ctx.check(self.ctx._lib.abcb200_chain_state(self._h, C.cast(C.byref(n), C.c_void_p), None, 0, None, None))
        th = np.empty((n.value, self.P), order="F"); w = np.empty(n.value); dv = np.empty(self.P)
        self.ctx.check(self.ctx._lib.abcb200_chain_state(self._h, None, _ptr(th), n.value, _ptr(w), _ptr(dv)))
        return th, w, dv

    def restore(self, theta, weights, dv, sets_done):
        th, w, d = _f(theta), _vec(weights), _vec(dv)
        if th.ndim != 2 or th.shape[1] != self.P or w.size != th.shape[0] or d.size != self.P:
            raise ValueError("restore: theta must be n x P, weights n, dv P")
        self.ctx.check(self.ctx._lib.abcb200_chain_restore(self._h, _ptr(th), th.shape[0], th.shape[0], _ptr(w), _ptr(d), int(sets_done)))

    def close(self):
        if self._h:
            self.ctx._lib.abcb200_chain_destroy(self._h)
            self._h = None
