// sample.cu — next-set proposal sampling, the step right after the hot path.
//
// Reference: ABC::sample_predictive_priors src/AbcUtil.cpp:378-390 = ABC::sample_posterior (:366-376; a weighted draw
// of rows through ABC::gsl_rng_nonuniform_int :111-121, i.e. gsl_ran_discrete: P(row j) = w_j / sum w) followed, per sampled
// row, by ABC::gsl_ran_trunc_normal (:146-158): every parameter gets Prior::noise (include/AbcSmc/Priors.h:18-41) —
// dev = recast(mu + sigma * N(0,1)) with sigma = sqrt(doubled variance), redrawn until valid(dev) (likelihood != 0,
// Parameter.h:77) at most MAX_ATTEMPTS = 1000 times, else the prior's mean (and a message on stderr).
//
// Parity is DISTRIBUTIONAL: the reference consumes one gsl_rng stream sequentially (Walker alias tables + polar
// gaussians); here every (sample, parameter, attempt) owns a Philox-4x32-10 counter, so the draw is reproducible from
// (seed) alone, independent of the launch geometry, and needs no state in memory. Rows are drawn by inverting the
// cumulative weights (exact same probabilities as the alias method).
//
// Data: weights n_pp, theta n_pp x P column-major, out num_samples x P column-major. O(num_samples * (log n_pp + P)).
#include "kernels.cuh"

namespace {

// Philox-4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0, k1)
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox(uint64_t seed, uint64_t ctr_lo, uint32_t ctr_a, uint32_t ctr_b, uint32_t (&out)[4]) {
    uint32_t c[4] = {(uint32_t)ctr_lo, (uint32_t)(ctr_lo >> 32), ctr_a, ctr_b};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = c[i];
}
// uniform in the OPEN interval (0, 1): 52 random bits + 1/2, every value exactly representable ((bits + 0.5) needs 53 bits), so
// neither 0 nor 1 can come out (with 53 bits the + 0.5 was a rounding tie for half the draws and the top value rounded to 1.0)
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
    const uint64_t bits = (((uint64_t)hi << 32) | lo) >> 12;
    return ((double)bits + 0.5) * (1.0 / 4503599627370496.0);
}

// Inclusive prefix sums of the weights, one CTA, fixed order (tile after tile with a running carry).
constexpr int CDF_T = 1024;
__global__ void __launch_bounds__(CDF_T) cdf_kernel(const double* __restrict__ w, int64_t n, double* __restrict__ cdf) {
    __shared__ double wsum[CDF_T / 32];
    __shared__ double s_carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_carry = 0.0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += CDF_T) {
        const int64_t i = base + tid;
        double v = (i < n) ? w[i] : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
        if (lane == 31) wsum[wid] = v;
        __syncthreads();
        double off = s_carry;
        for (int ww = 0; ww < wid; ww++) off += wsum[ww];
        if (i < n) cdf[i] = off + v;
        __syncthreads();
        if (tid == CDF_T - 1) s_carry = off + v;
        __syncthreads();
    }
}

struct SampleArgs {
    uint64_t seed;
    int64_t num_samples, n_pp, ld, ld_out;
    int P, max_attempts;
    const double* cdf; const double* theta; const double* dv;
    const double* lo; const double* hi; const int32_t* integral; const double* prior_mean;
    double* out; uint64_t* parent; unsigned long long* fallbacks;
};

__global__ void __launch_bounds__(256) sample_kernel(SampleArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.num_samples) return;
    uint32_t r[4];
    // ---- weighted draw of the parent row (AbcUtil.cpp:111-121, 366-376) ------------------------------------------------
    philox(a.seed, (uint64_t)i, 0u, 0u, r);
    const double total = a.cdf[a.n_pp - 1];
    const double x = u53(r[0], r[1]) * total;
    int64_t lo = 0, hi = a.n_pp - 1;                 // first j with cdf[j] > x (zero-weight rows are never drawn)
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a.cdf[mid] > x) hi = mid; else lo = mid + 1;
    }
    const int64_t j = lo;
    if (a.parent) a.parent[i] = (uint64_t)j;
    // ---- independent truncated normal noise per parameter (AbcUtil.cpp:146-158, Priors.h:18-41) ---------------------
    for (int p = 0; p < a.P; p++) {
        const double mu = a.theta[(int64_t)p * a.ld + j];
        const double sigma = sqrt(a.dv[p]);
        const double lop = a.lo[p], hip = a.hi[p];
        const bool integral = a.integral && a.integral[p] != 0;
        double dev = 0.0;
        bool ok = false;
        for (int att = 0; att < a.max_attempts && !ok; att += 2) {      // one Philox block = one Box-Muller pair = two attempts
            philox(a.seed, (uint64_t)i, (uint32_t)(p + 1), (uint32_t)(att >> 1), r);
            const double u1 = u53(r[0], r[1]), u2 = u53(r[2], r[3]);
            const double rad = sqrt(-2.0 * log(u1));
            double sn, cs;
            sincospi(2.0 * u2, &sn, &cs);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                if (ok || att + h >= a.max_attempts) continue;
                double v = fma(sigma, rad * (h == 0 ? cs : sn), mu);
                if (integral) v = round(v);                             // DiscreteUniformPrior::recast (Priors.h:80)
                if (v >= lop && v <= hip) { dev = v; ok = true; }       // valid(): likelihood != 0 (Priors.h:76-78, 101-103)
            }
        }
        if (!ok) {                                                      // Priors.h:26-28: fall back to the prior's mean
            dev = a.prior_mean[p];
            if (a.fallbacks) atomicAdd(a.fallbacks, 1ull);
        }
        a.out[(int64_t)p * a.ld_out + i] = dev;
    }
}

// ---- multivariate noise (NOISE::MULTIVARIATE, AbcSmc.cpp:491-503) ---------------------------------------------------------
// ABC::setup_mvn_sampler (src/AbcUtil.cpp:462-488): sample variance-covariance matrix of the predictive prior's rows
// (gsl_ran_multivariate_gaussian_vcov: sum (x_i - mean)(x_i - mean)^T / (n - 1), GSL manual), diagonal doubled (:477-480),
// then its Cholesky factor L (gsl_linalg_cholesky_decomp1). ABC::sample_mvn_predictive_priors (:392-404) draws the parent row
// like the independent variant and adds L z, z ~ N(0, I) (gsl_ran_multivariate_gaussian), redrawing the WHOLE vector until every
// parameter is valid after recast, parameters checked in order and the check stopped at the first invalid one
// (gsl_ran_trunc_mv_normal, :123-144). The reference loops without a limit; here max_attempts bounds it and the samples that
// ran out keep the recast parent row and are counted (a deviation only where the reference would never return).
__global__ void __launch_bounds__(256) cov_kernel(const double* __restrict__ X, int64_t ld, int64_t n, int P, const double* __restrict__ mean,
                                                  double* __restrict__ S) {
    __shared__ double red[32];
    const int j = blockIdx.x, k = blockIdx.y;
    if (k > j) return;
    const double mj = mean[j], mk = mean[k];
    const double* xj = X + (int64_t)j * ld;
    const double* xk = X + (int64_t)k * ld;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 256) s = fma(xj[i] - mj, xk[i] - mk, s);
    s = block_sum(s, red);
    if (threadIdx.x == 0) {
        double v = s / ((double)n - 1.0);
        if (j == k) v *= 2.0;                                   // AbcUtil.cpp:477-480
        S[(int64_t)k * P + j] = v; S[(int64_t)j * P + k] = v;
    }
}
// In-place lower Cholesky factor of the P x P matrix S (column-major, ld P), one CTA; flag[0] = 1 when S is not positive definite.
__global__ void __launch_bounds__(128) chol_kernel(double* __restrict__ S, int P, int* __restrict__ flag) {
    extern __shared__ double A[];
    const int tid = threadIdx.x;
    for (int e = tid; e < P * P; e += 128) A[e] = S[e];
    __syncthreads();
    for (int c = 0; c < P; c++) {
        const double d = A[c * P + c];
        if (!(d > 0.0)) { if (tid == 0) flag[0] = 1; return; }   // uniform: every thread reads the same d
        const double r = sqrt(d);
        __syncthreads();
        for (int i = c + tid; i < P; i += 128) A[c * P + i] = (i == c) ? r : A[c * P + i] / r;
        __syncthreads();
        for (int e = tid; e < (P - c - 1) * (P - c - 1); e += 128) {
            const int jj = c + 1 + e / (P - c - 1), ii = c + 1 + e % (P - c - 1);
            if (ii >= jj) A[jj * P + ii] = fma(-A[c * P + ii], A[c * P + jj], A[jj * P + ii]);
        }
        __syncthreads();
    }
    for (int e = tid; e < P * P; e += 128) { const int jj = e / P, ii = e % P; S[e] = (ii >= jj) ? A[e] : 0.0; }
}

constexpr int MVN_PMAX = 128;
struct MvnArgs {
    uint64_t seed;
    int64_t num_samples, n_pp, ld, ld_out;
    int P, max_attempts;
    const double* cdf; const double* theta; const double* L;
    const double* lo; const double* hi; const int32_t* integral;
    double* out; uint64_t* parent; unsigned long long* failures;
};
__global__ void __launch_bounds__(128) sample_mvn_kernel(MvnArgs a) {
    extern __shared__ double Ls[];                 // L packed by rows: row p at p (p + 1) / 2
    const int P = a.P;
    for (int e = threadIdx.x; e < P * (P + 1) / 2; e += blockDim.x) {
        int p = 0; while ((p + 1) * (p + 2) / 2 <= e) p++;
        const int q = e - p * (p + 1) / 2;
        Ls[e] = a.L[(int64_t)q * P + p];
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.num_samples) return;
    uint32_t r[4];
    philox(a.seed, (uint64_t)i, 0u, 0u, r);        // same parent draw as the independent variant
    const double total = a.cdf[a.n_pp - 1];
    const double x = u53(r[0], r[1]) * total;
    int64_t lo = 0, hi = a.n_pp - 1;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (a.cdf[mid] > x) hi = mid; else lo = mid + 1; }
    const int64_t j = lo;
    if (a.parent) a.parent[i] = (uint64_t)j;
    double z[MVN_PMAX], v[MVN_PMAX];
    bool ok = false;
    for (int att = 0; att < a.max_attempts && !ok; att++) {
        for (int b = 0; b < (P + 1) / 2; b++) {    // P standard normals: one Philox block per Box-Muller pair
            philox(a.seed, (uint64_t)i, 0x80000000u | (uint32_t)b, (uint32_t)att, r);
            const double rad = sqrt(-2.0 * log(u53(r[0], r[1])));
            double sn, cs;
            sincospi(2.0 * u53(r[2], r[3]), &sn, &cs);
            z[2 * b] = rad * cs;
            if (2 * b + 1 < P) z[2 * b + 1] = rad * sn;
        }
        ok = true;
        for (int p = 0; p < P && ok; p++) {        // AbcUtil.cpp:136-139: stop at the first invalid parameter
            double acc = a.theta[(int64_t)p * a.ld + j];
            const double* lp = Ls + p * (p + 1) / 2;
            for (int q = 0; q <= p; q++) acc = fma(lp[q], z[q], acc);
            if (a.integral && a.integral[p]) acc = round(acc);
            v[p] = acc;
            ok = (acc >= a.lo[p]) && (acc <= a.hi[p]);
        }
    }
    if (!ok) {
        for (int p = 0; p < P; p++) { double m = a.theta[(int64_t)p * a.ld + j]; if (a.integral && a.integral[p]) m = round(m); v[p] = m; }
        if (a.failures) atomicAdd(a.failures, 1ull);
    }
    for (int p = 0; p < P; p++) a.out[(int64_t)p * a.ld_out + i] = v[p];
}

}  // namespace

size_t sample_ws_bytes(int64_t n_pp) { return align_up((size_t)n_pp * 8, 256) + 512; }
size_t mvn_setup_ws_bytes(int64_t n_pp, int P) { return moments_ws_bytes(n_pp, P) + align_up((size_t)P * 8, 256) + 1024; }

// All pointers are device pointers; fallbacks (nullable) must be zeroed by the caller.
int sample_predictive_priors_core(abcb200_ctx* ctx, uint64_t seed, int64_t num_samples, const double* weights, const double* theta,
                                  int64_t ld, int64_t n_pp, int P, const double* dv, const double* lo, const double* hi,
                                  const int32_t* integral, const double* prior_mean, int max_attempts, double* out, int64_t ld_out,
                                  uint64_t* parent, unsigned long long* fallbacks) {
    double* cdf = ws_new<double>(ctx, (size_t)n_pp);
    if (!cdf) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in sample_predictive_priors");
    LAUNCH(ctx, cdf_kernel, 1, CDF_T, 0, weights, n_pp, cdf);
    SampleArgs a;
    a.seed = seed; a.num_samples = num_samples; a.n_pp = n_pp; a.ld = ld; a.ld_out = ld_out; a.P = P; a.max_attempts = max_attempts;
    a.cdf = cdf; a.theta = theta; a.dv = dv; a.lo = lo; a.hi = hi; a.integral = integral; a.prior_mean = prior_mean;
    a.out = out; a.parent = parent; a.fallbacks = fallbacks;
    LAUNCH(ctx, sample_kernel, (unsigned)((num_samples + 255) / 256), 256, 0, a);
    return ABCB200_OK;
}

// L (P x P, column-major, ld P, lower triangle; device) from theta (n_pp x P, device); flag (device int, zeroed by the caller) is
// set when the doubled-diagonal covariance is not positive definite (gsl_linalg_cholesky_decomp1 would raise GSL_EDOM).
int setup_mvn_sampler_core(abcb200_ctx* ctx, const double* theta, int64_t ld, int64_t n_pp, int P, double* L, int* flag) {
    int nchunk = 0;
    double* stats = (double*)ws_alloc(ctx, moments_ws_bytes(n_pp, P));
    double* mean = ws_new<double>(ctx, P);
    if (!stats || !mean) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in setup_mvn_sampler");
    ABC_TRY(launch_col_stats(ctx, theta, ld, n_pp, P, stats, &nchunk));
    ABC_TRY(launch_col_finalize(ctx, stats, nchunk, n_pp, P, mean, nullptr, 1.0, nullptr));
    LAUNCH(ctx, cov_kernel, dim3(P, P), 256, 0, theta, ld, n_pp, P, mean, L);
    const size_t smem = (size_t)P * P * sizeof(double);
    CUDA_TRY(ctx, cudaFuncSetAttribute(chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAUNCH(ctx, chol_kernel, 1, 128, smem, L, P, flag);
    return ABCB200_OK;
}

int sample_mvn_predictive_priors_core(abcb200_ctx* ctx, uint64_t seed, int64_t num_samples, const double* weights, const double* theta,
                                      int64_t ld, int64_t n_pp, int P, const double* L, const double* lo, const double* hi,
                                      const int32_t* integral, int max_attempts, double* out, int64_t ld_out, uint64_t* parent,
                                      unsigned long long* failures) {
    if (P > MVN_PMAX) ABC_FAIL(ctx, ABCB200_EINVAL, "sample_mvn_predictive_priors: P=%d exceeds %d", P, MVN_PMAX);
    double* cdf = ws_new<double>(ctx, (size_t)n_pp);
    if (!cdf) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in sample_mvn_predictive_priors");
    LAUNCH(ctx, cdf_kernel, 1, CDF_T, 0, weights, n_pp, cdf);
    MvnArgs a;
    a.seed = seed; a.num_samples = num_samples; a.n_pp = n_pp; a.ld = ld; a.ld_out = ld_out; a.P = P; a.max_attempts = max_attempts;
    a.cdf = cdf; a.theta = theta; a.L = L; a.lo = lo; a.hi = hi; a.integral = integral; a.out = out; a.parent = parent; a.failures = failures;
    const size_t smem = (size_t)P * (P + 1) / 2 * sizeof(double);
    CUDA_TRY(ctx, cudaFuncSetAttribute(sample_mvn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAUNCH(ctx, sample_mvn_kernel, (unsigned)((num_samples + 127) / 128), 128, smem, a);
    return ABCB200_OK;
}
