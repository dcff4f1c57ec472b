// chain.cu — one call per SMC set with the previous set's predictive prior RESIDENT on the device.
//
// Reference: the body of the set loop of AbcSmc::read_SMC_sets_from_database, src/AbcSmc.cpp:634-664 — rank the set's particles
// (particle_ranking_PLS / _simple), keep the first next_pred_prior_size, gather their parameters and metrics (:648-649), report on them
// (AbcLog::filtering_report, src/AbcLog.cpp:81-124: ABC::calculate_nrmse src/AbcUtil.cpp:326-345, column means, ABC::median :46-61) — and
// AbcSmc::calculate_predictive_prior_weights, :1041-1066 (doubled variance of the gathered rows, weights against set t-1's gathered rows,
// weights and doubled variance). The reference keeps every set's _weights / _doubled_variance on the host and RECOMPUTES them for all
// earlier sets on every `--process` run (the call sits inside the set loop, :664). Here set t-1's gathered parameters, weights and doubled
// variance stay in device buffers owned by the chain between calls: nothing of the previous set is uploaded or evaluated again, and a host
// that persisted them (abcb200_chain_state after each set) re-seeds a new process with abcb200_chain_restore instead of replaying the sets.
#include <cmath>
#include <new>

#include "kernels.cuh"

struct abcb200_chain {
    abcb200_ctx* ctx;
    int P;
    int sets;                 // sets processed (or restored) so far
    int cur;                  // buffers of the last finished set
    double* theta[2];         // n x P gathered parameters (ld n), rank order
    double* w[2];             // n weights
    double* dv[2];            // P doubled variances
    int64_t n[2], cap[2];
};

namespace {

inline int64_t pad32(int64_t n) { return (n + 31) / 32 * 32; }

__device__ __forceinline__ uint64_t sort_key(double x) {      // monotone map double -> uint64 (as sort.cu's order_key)
    const uint64_t b = (x == 0.0) ? 0ull : (uint64_t)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_value(uint64_t k) {
    const uint64_t b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
// keys of the columns of two matrices laid side by side: segment c < Pa from A (n x Pa, lda), the rest from B
__global__ void column_keys_kernel(const double* __restrict__ A, int64_t lda, int Pa, const double* __restrict__ B, int64_t ldb, int64_t n,
                                   uint64_t* __restrict__ keys) {
    const int c = blockIdx.y;
    const double* src = (c < Pa) ? A + (int64_t)c * lda : B + (int64_t)(c - Pa) * ldb;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) keys[(int64_t)c * n + i] = sort_key(src[i]);
}
// ABC::median (src/AbcUtil.cpp:46-61) from sorted segments
__global__ void median_kernel(const uint64_t* __restrict__ keys, int64_t n, int ncol, double* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncol) return;
    const uint64_t* k = keys + (int64_t)c * n;
    out[c] = (n % 2 == 0) ? (key_value(k[n / 2 - 1]) + key_value(k[n / 2])) / 2 : key_value(k[n / 2]);
}
// numer[i] = prod_p prior_p.likelihood(theta[i, p]) (src/AbcUtil.cpp:559-561) for the three prior kinds of include/AbcSmc/Priors.h:
// type 0 ContinuousUniformPrior [a, b] (:101-103), 1 DiscreteUniformPrior [a, b] (:75-77), 2 GaussianPrior mean a, sd b (:53-55, gsl_ran_gaussian_pdf)
__global__ void prior_numer_kernel(const double* __restrict__ th, int64_t ld, int64_t n, int P, const int32_t* __restrict__ type, const double* __restrict__ a,
                                   const double* __restrict__ b, double* __restrict__ numer) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double num = 1.0;
        for (int p = 0; p < P; p++) {
            const double v = th[(int64_t)p * ld + i];
            double l;
            if (type[p] == 0) l = (a[p] <= v && v <= b[p]) ? 1.0 / (b[p] - a[p]) : 0.0;
            else if (type[p] == 1) { const double mn = (double)(long long)a[p], mx = (double)(long long)b[p]; l = (v == round(v) && mn <= v && v <= mx) ? 1.0 / (mx - mn + 1.0) : 0.0; }
            else { const double u = (v - a[p]) / fabs(b[p]); l = (1.0 / (sqrt(2.0 * M_PI) * fabs(b[p]))) * exp(-u * u / 2.0); }
            num *= l;
        }
        numer[i] = num;
    }
}

int chain_reserve(abcb200_chain* ch, int slot, int64_t n) {
    abcb200_ctx* ctx = ch->ctx;
    if (n <= ch->cap[slot]) return ABCB200_OK;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (ch->theta[slot]) cudaFree(ch->theta[slot]);
    if (ch->w[slot]) cudaFree(ch->w[slot]);
    ch->theta[slot] = ch->w[slot] = nullptr; ch->cap[slot] = 0;
    if (cudaMalloc(&ch->theta[slot], (size_t)n * ch->P * 8) != cudaSuccess || cudaMalloc(&ch->w[slot], (size_t)n * 8) != cudaSuccess) {
        cudaGetLastError();
        ABC_FAIL(ctx, ABCB200_ENOMEM, "chain: device buffers for %lld x %d predictive-prior rows", (long long)n, ch->P);
    }
    ch->cap[slot] = n;
    return ABCB200_OK;
}

}  // namespace

extern "C" int abcb200_chain_create(abcb200_ctx* ctx, int P, abcb200_chain** out) {
    if (!out) return ABCB200_EINVAL;
    *out = nullptr;
    if (!ctx || P < 1 || P > 128) return ABCB200_EINVAL;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return ABCB200_ENODEV;
    abcb200_chain* ch = new (std::nothrow) abcb200_chain();
    if (!ch) return ABCB200_ENOMEM;
    memset(ch, 0, sizeof(*ch));
    ch->ctx = ctx; ch->P = P;
    for (int s = 0; s < 2; s++)
        if (cudaMalloc(&ch->dv[s], (size_t)P * 8) != cudaSuccess) { cudaGetLastError(); if (ch->dv[0]) cudaFree(ch->dv[0]); delete ch; return ABCB200_ENOMEM; }
    *out = ch;
    return ABCB200_OK;
}

extern "C" int abcb200_chain_destroy(abcb200_chain* ch) {
    if (!ch) return ABCB200_OK;
    cudaSetDevice(ch->ctx->device);
    cudaStreamSynchronize(ch->ctx->stream);
    for (int s = 0; s < 2; s++) { if (ch->theta[s]) cudaFree(ch->theta[s]); if (ch->w[s]) cudaFree(ch->w[s]); if (ch->dv[s]) cudaFree(ch->dv[s]); }
    delete ch;
    return ABCB200_OK;
}

extern "C" int abcb200_chain_sets(const abcb200_chain* ch) { return ch ? ch->sets : 0; }
extern "C" int abcb200_chain_nparams(const abcb200_chain* ch) { return ch ? ch->P : 0; }

// Re-seed the chain with a finished set that the host persisted (its gathered parameters in rank order, weights, doubled variance).
extern "C" int abcb200_chain_restore(abcb200_chain* ch, const double* theta, int64_t ld, int64_t n, const double* weights, const double* dv, int sets_done) {
    if (!ch) return ABCB200_EINVAL;
    abcb200_ctx* ctx = ch->ctx;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return ABCB200_ENODEV;
    if (!theta || !weights || !dv || n < 1 || ld < n || sets_done < 1) ABC_FAIL(ctx, ABCB200_EINVAL, "chain_restore: bad argument");
    const int slot = ch->cur ^ 1;
    ABC_TRY(chain_reserve(ch, slot, n));
    CUDA_TRY(ctx, cudaMemcpy2DAsync(ch->theta[slot], (size_t)n * 8, theta, (size_t)ld * 8, (size_t)n * 8, (size_t)ch->P, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ch->w[slot], weights, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ch->dv[slot], dv, (size_t)ch->P * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ch->n[slot] = n; ch->cur = slot; ch->sets = sets_done;
    return ABCB200_OK;
}

// The last finished set as the device holds it (what a host persists): theta_out n x P (ld ld_out), weights_out n, dv_out P; all nullable.
extern "C" int abcb200_chain_state(abcb200_chain* ch, int64_t* n_out, double* theta_out, int64_t ld_out, double* weights_out, double* dv_out) {
    if (!ch) return ABCB200_EINVAL;
    abcb200_ctx* ctx = ch->ctx;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return ABCB200_ENODEV;
    if (ch->sets < 1) ABC_FAIL(ctx, ABCB200_EINVAL, "chain_state: no set has been processed yet");
    const int s = ch->cur;
    const int64_t n = ch->n[s];
    if (n_out) *n_out = n;
    if (theta_out) {
        if (ld_out < n) ABC_FAIL(ctx, ABCB200_EINVAL, "chain_state: leading dimension < rows");
        CUDA_TRY(ctx, cudaMemcpy2DAsync(theta_out, (size_t)ld_out * 8, ch->theta[s], (size_t)n * 8, (size_t)n * 8, (size_t)ch->P, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (weights_out) CUDA_TRY(ctx, cudaMemcpyAsync(weights_out, ch->w[s], (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (dv_out) CUDA_TRY(ctx, cudaMemcpyAsync(dv_out, ch->dv[s], (size_t)ch->P * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ABCB200_OK;
}

extern "C" int abcb200_chain_process_set(abcb200_chain* ch, const double* met, int64_t ld_met, const double* par, int64_t ld_par, int64_t N, int K,
                                         const double* target, int filter, double training_fraction, int method, int64_t top_n,
                                         const int32_t* prior_type, const double* prior_a, const double* prior_b, const double* numer_all,
                                         uint64_t* order_out, double* weights_out, double* dv_out, double* report_out, int* n_comp_used_out) {
    if (!ch) return ABCB200_EINVAL;
    abcb200_ctx* ctx = ch->ctx;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return ABCB200_ENODEV;
    ctx->err[0] = 0;
    const int P = ch->P;
    const bool simple = filter == 1;
    if (!met || !par || !target || !order_out || (filter != 0 && filter != 1)) ABC_FAIL(ctx, ABCB200_EINVAL, "chain_process_set: bad argument");
    if (ld_met < N || ld_par < N) ABC_FAIL(ctx, ABCB200_EINVAL, "chain_process_set: leading dimension < N");
    if ((prior_type != nullptr) != (prior_a != nullptr) || (prior_type != nullptr) != (prior_b != nullptr)) ABC_FAIL(ctx, ABCB200_EINVAL, "chain_process_set: prior_type / prior_a / prior_b go together");
    ABC_TRY(rank_shape_check(ctx, N, K, P, training_fraction, method, simple));
    if (top_n <= 0 || top_n > N) top_n = N;
    const int64_t n = top_n, ldd = pad32(N);
    const int prev = ch->cur, nxt = (ch->sets > 0) ? (ch->cur ^ 1) : ch->cur;
    ABC_TRY(chain_reserve(ch, nxt, n));
    const int ncol = P + K;
    size_t need = rank_ws_bytes(ctx, N, K, P, training_fraction, method, simple);
    need += align_up((size_t)ldd * K * 8, 256) + align_up((size_t)ldd * P * 8, 256) + align_up((size_t)K * 8, 256) + align_up((size_t)N * 8, 256);   // staged inputs, order
    need += align_up((size_t)n * K * 8, 256) + moments_ws_bytes(n, P) + moments_ws_bytes(n, K) + 4 * align_up((size_t)ncol * 8, 256);                  // gathered metrics, stats
    need += 2 * align_up((size_t)ncol * n * 8, 256) + radix_hist_bytes(n, ncol);                                                                        // medians
    need += 3 * align_up((size_t)N * 8, 256) + 3 * align_up((size_t)P * 8, 256);                                                                        // numerators, flat priors, distances (tie order 1)
    if (ch->sets > 0) need += weights_ws_bytes(ctx, n, ch->n[prev], P);
    ABC_TRY(ws_reserve(ctx, need + 16384));
    double* d_met = ws_new<double>(ctx, (size_t)ldd * K);
    double* d_par = ws_new<double>(ctx, (size_t)ldd * P);
    double* d_target = ws_new<double>(ctx, K);
    uint64_t* d_order = ws_new<uint64_t>(ctx, N);
    double* Gm = ws_new<double>(ctx, (size_t)n * K);
    double* rep = ws_new<double>(ctx, 2 * (size_t)ncol);          // means (P + K), medians (P + K)
    uint64_t* keys = ws_new<uint64_t>(ctx, (size_t)ncol * n);
    uint64_t* keys_alt = ws_new<uint64_t>(ctx, (size_t)ncol * n);
    uint32_t* hist = (uint32_t*)ws_alloc(ctx, radix_hist_bytes(n, ncol));
    double* d_numer_all = numer_all ? ws_new<double>(ctx, N) : nullptr;
    double* d_numer = (numer_all || prior_type) ? ws_new<double>(ctx, n) : nullptr;
    int32_t* d_ptype = prior_type ? ws_new<int32_t>(ctx, P) : nullptr;
    double* d_pa = prior_type ? ws_new<double>(ctx, P) : nullptr;
    double* d_pb = prior_type ? ws_new<double>(ctx, P) : nullptr;
    double* d_ss = ws_new<double>(ctx, 1);
    double* d_dist = (ctx->tie_order == 1) ? ws_new<double>(ctx, N) : nullptr;
    if (ctx->tie_order == 1 && !d_dist) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in chain_process_set");
    if (!d_met || !d_par || !d_target || !d_order || !Gm || !rep || !keys || !keys_alt || !hist || !d_ss || (numer_all && !d_numer_all) ||
        ((numer_all || prior_type) && !d_numer) || (prior_type && (!d_ptype || !d_pa || !d_pb)))
        ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in chain_process_set");
    Arrival arr;
    ABC_TRY(stage_inputs(ctx, d_met, d_par, d_target, ldd, met, ld_met, par, ld_par, target, N, K, P, &arr));      // column blocks on the copy stream
    stage_begin(ctx, 8);
    if (numer_all) CUDA_TRY(ctx, cudaMemcpyAsync(d_numer_all, numer_all, sizeof(double) * N, cudaMemcpyHostToDevice, ctx->stream));
    if (prior_type) {
        CUDA_TRY(ctx, cudaMemcpyAsync(d_ptype, prior_type, sizeof(int32_t) * P, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(d_pa, prior_a, sizeof(double) * P, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(d_pb, prior_b, sizeof(double) * P, cudaMemcpyHostToDevice, ctx->stream));
    }
    stage_end(ctx, 8);
    // ---- filtering (AbcSmc.cpp:634-646) --------------------------------------------------------------------------------------
    ABC_TRY(rank_on_device(ctx, d_met, ldd, d_par, ldd, N, K, P, d_target, training_fraction, method, n, d_order, d_dist, n_comp_used_out, nullptr, simple, &arr));
    if (d_dist) ABC_TRY(tie_order_stdsort_device(ctx, d_dist, N, n, d_order));      // exact ties as std::sort leaves them (abcb200_set_tie_order)
    // ---- posterior rows (:648-649), their report statistics and the doubled variance (:1042-1047) -----------------------------
    double* th = ch->theta[nxt];
    ABC_TRY(launch_gather_rows(ctx, d_par, ldd, d_order, n, P, th, n));
    ABC_TRY(launch_gather_rows(ctx, d_met, ldd, d_order, n, K, Gm, n));
    stage_begin(ctx, 6);
    int nchunk = 0;
    double* stats_p = (double*)ws_alloc(ctx, moments_ws_bytes(n, P));
    double* stats_m = (double*)ws_alloc(ctx, moments_ws_bytes(n, K));
    if (!stats_p || !stats_m) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in chain_process_set");
    ABC_TRY(launch_col_stats(ctx, th, n, n, P, stats_p, &nchunk));
    ABC_TRY(launch_col_finalize(ctx, stats_p, nchunk, n, P, rep, nullptr, 2.0, ch->dv[nxt]));        // means + 2 * sample variance (AbcUtil.cpp:534)
    ABC_TRY(launch_col_stats(ctx, Gm, n, n, K, stats_m, &nchunk));
    ABC_TRY(launch_col_finalize(ctx, stats_m, nchunk, n, K, rep + P, nullptr, 1.0, nullptr));
    stage_end(ctx, 6);
    if (report_out) {
        const int gx = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 64));
        LAUNCH(ctx, column_keys_kernel, dim3(gx, ncol), 256, 0, th, n, P, Gm, n, n, keys);
        ABC_TRY(radix_sort_segments(ctx, keys, keys_alt, nullptr, nullptr, n, ncol, hist, nullptr));
        LAUNCH(ctx, median_kernel, (ncol + 127) / 128, 128, 0, keys, n, ncol, rep + ncol);
    }
    // ---- weights (:1049-1064) ---------------------------------------------------------------------------------------------------
    stage_begin(ctx, 7);
    if (ch->sets == 0) {
        ABC_TRY(launch_fill(ctx, ch->w[nxt], n, 1.0 / (double)n));                                      // AbcUtil.cpp:543-544
    } else {
        if (numer_all) ABC_TRY(launch_gather_rows(ctx, d_numer_all, N, d_order, n, 1, d_numer, n));
        else if (prior_type) {
            const int gx = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)4 * ctx->sm_count));
            LAUNCH(ctx, prior_numer_kernel, gx, 256, 0, th, n, n, P, d_ptype, d_pa, d_pb, d_numer);
        }
        ABC_TRY(weights_unnorm_dev(ctx, d_numer, th, n, n, ch->theta[prev], ch->n[prev], ch->n[prev], ch->w[prev], ch->dv[prev], P, 0, ch->w[nxt], d_ss));
        ABC_TRY(launch_scale_weights(ctx, ch->w[nxt], n, d_ss));
    }
    stage_end(ctx, 7);
    stage_begin(ctx, 9);
    ABC_TRY(hpin_reserve(ctx, sizeof(double) * 2 * (size_t)ncol + 64));
    double* h_rep = (double*)ctx->hpin;
    CUDA_TRY(ctx, cudaMemcpyAsync(order_out, d_order, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (weights_out) CUDA_TRY(ctx, cudaMemcpyAsync(weights_out, ch->w[nxt], sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (dv_out) CUDA_TRY(ctx, cudaMemcpyAsync(dv_out, ch->dv[nxt], sizeof(double) * (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
    if (report_out) CUDA_TRY(ctx, cudaMemcpyAsync(h_rep, rep, sizeof(double) * 2 * (size_t)ncol, cudaMemcpyDeviceToHost, ctx->stream));
    stage_end(ctx, 9);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (report_out) {
        // report_out: [0] NRMSE, then P + K posterior means, then P + K posterior medians (parameters first, as filtering_report prints them)
        const double* mean_met = h_rep + P;
        double acc = 0.0;                                                   // ABC::calculate_nrmse, src/AbcUtil.cpp:326-345
        for (int k = 0; k < K; k++) {
            const double sim = mean_met[k], obs = target[k];
            double expected = (std::fabs(obs) + std::fabs(sim)) / 2.0;
            if (sim == obs) expected = 1.0;
            const double d = (sim - obs) / expected;
            acc += d * d;
        }
        report_out[0] = std::sqrt(acc / (double)K);
        for (int c = 0; c < 2 * ncol; c++) report_out[1 + c] = h_rep[c];
    }
    ch->n[nxt] = n; ch->cur = nxt; ch->sets += 1;
    return ABCB200_OK;
}
