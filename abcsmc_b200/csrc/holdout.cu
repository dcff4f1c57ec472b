// holdout.cu — S3 hold-out residuals + PRESS and S4 Wilcoxon component selection.
//
// Reference: Model::cv_NEW_DATA lib/PLS/src/pls.cpp:494-510, PLS::validation :235-261,
//            PLS::optimal_num_components :265-289, PLS::wilcoxon :190-211, PLS::normalcdf :152-160.
//
// The reference materialises an M x n_te x A error cube (4.5 GB at the dengue config) from A dense products
// Y - X (R_c Q_c^T). Here the hold-out scores T = X_te R are formed once (DMMA, launch_xb) and residuals are the
// running prefix e_c = e_{c-1} - t_c q_c^T, so only T (n_te x A) lives in HBM. PRESS comes from one pass over T.
// For the selection, the signed-rank tests of a round (every undecided response y, a chunk of consecutive
// `alt` component counts) are generated as sortable 64-bit keys, sorted together by the segmented radix sort
// (sort.cu) and reduced to the exact integer rank sums. Keys: bits(|d|) << 1 | (d > 0), d = |e_ref| - |e_alt|;
// the sign of d rides in the LSB, d == 0 gives key 0 and sign 0 (pls.cpp:196). Exact ties in |d| between opposite
// signs are ordered negative-first (the reference's order there is whatever introsort yields).
//
// Screening: only the DECISION p > alpha is consumed (pls.cpp:283), and p is monotone in the rank sum d. One CTA per
// test bins the keys with a monotone map into 4096 buckets (signed counts in shared memory), which brackets every
// element's rank to its bucket and therefore d to an exact integer interval [d_lo, d_hi] (positives at the bottom /
// top of each bucket). If p(d_lo) > alpha the test certainly succeeds, if p(d_hi) <= alpha it certainly fails; only
// tests whose interval straddles the threshold are sorted exactly. Decisions are therefore identical to sorting
// every test, at 8 B/key of traffic instead of 192 B/key.
#include "kernels.cuh"

namespace {

constexpr int PR_THREADS = 256;
constexpr int PR_ROWS = 2;     // rows per thread

// PRESS partials: for each component c, e[y] -= T[i,c] * Q[y,c]; press[y,c] += e[y]^2 over the CTA's rows.
// partial[cta][y*A + c]
template <int MCHUNK>
__global__ void __launch_bounds__(PR_THREADS) press_kernel(const double* __restrict__ T, int64_t ldt, const double* __restrict__ Y,
                                                           int64_t ldy, int64_t n, int M, int A, const double* __restrict__ Q,
                                                           int y0, double* __restrict__ partial) {
    __shared__ double wsum[PR_THREADS / 32][MCHUNK];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int mc = min(MCHUNK, M - y0);
    double* out = partial + (int64_t)blockIdx.x * M * A;
    for (int i = tid; i < mc * A; i += PR_THREADS) { const int y = i / A, c = i - y * A; out[(int64_t)(y0 + y) * A + c] = 0.0; }
    __syncthreads();
    const int64_t rows_per_blk = (int64_t)PR_THREADS * PR_ROWS;
    const int64_t nblk = (n + rows_per_blk - 1) / rows_per_blk;
    for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const int64_t i0 = blk * rows_per_blk + tid, i1 = i0 + PR_THREADS;
        const bool v0 = i0 < n, v1 = i1 < n;
        double e0[MCHUNK], e1[MCHUNK];
#pragma unroll
        for (int y = 0; y < MCHUNK; y++) {
            e0[y] = (v0 && y < mc) ? Y[(int64_t)(y0 + y) * ldy + i0] : 0.0;
            e1[y] = (v1 && y < mc) ? Y[(int64_t)(y0 + y) * ldy + i1] : 0.0;
        }
        for (int c = 0; c < A; c++) {
            const double t0 = v0 ? T[(int64_t)c * ldt + i0] : 0.0;
            const double t1 = v1 ? T[(int64_t)c * ldt + i1] : 0.0;
            const double* qc = Q + (int64_t)c * M + y0;
#pragma unroll
            for (int y = 0; y < MCHUNK; y++) {
                if (y < mc) {
                    const double q = qc[y];
                    e0[y] = fma(-t0, q, e0[y]);
                    e1[y] = fma(-t1, q, e1[y]);
                    double s = fma(e0[y], e0[y], e1[y] * e1[y]);
                    s = warp_sum(s);
                    if (lane == 0) wsum[wid][y] = s;
                }
            }
            __syncthreads();
            if (tid < mc) {
                double s = 0;
#pragma unroll
                for (int w = 0; w < PR_THREADS / 32; w++) s += wsum[w][tid];
                out[(int64_t)(y0 + tid) * A + c] += s;
            }
            __syncthreads();
        }
    }
}

// press[y + c*M] (M x A column-major) = sum over CTAs (fixed order); optional MSE scaling
__global__ void press_reduce_kernel(const double* __restrict__ partial, int ncta, int M, int A, double scale, double* __restrict__ press) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * A) return;
    const int y = i / A, c = i - y * A;
    double s = 0;
    for (int b = 0; b < ncta; b++) s += partial[(int64_t)b * M * A + i];
    press[(int64_t)c * M + y] = s * scale;
}

// ref[y] = first argmin_c press[y,c]  (Eigen minCoeff(&idx), pls.cpp:278)
__global__ void argmin_kernel(const double* __restrict__ press, int M, int A, int* __restrict__ ref) {
    const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (y >= M) return;
    const int lane = threadIdx.x & 31;
    double best = 0; int bi = -1;
    for (int c = lane; c < A; c += 32) {
        const double v = press[(int64_t)c * M + y];
        if (bi < 0 || v < best) { best = v; bi = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi >= 0 && (bi < 0 || ov < best || (ov == best && oi < bi))) { best = ov; bi = oi; }
    }
    if (lane == 0) ref[y] = bi;
}

// Eref[i,y] = residual with ref[y]+1 components; Ecur[i,y] = Y[i,y] (zero components)
__global__ void eref_kernel(const double* __restrict__ T, int64_t ldt, const double* __restrict__ Y, int64_t ldy, int64_t n, int M,
                            const double* __restrict__ Q, const int* __restrict__ ref, double* __restrict__ Eref,
                            double* __restrict__ Ecur) {
    const int y = blockIdx.y;
    const int nc = ref[y] + 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double y0 = Y[(int64_t)y * ldy + i];
        double e = y0;
        for (int c = 0; c < nc; c++) e = fma(-T[(int64_t)c * ldt + i], Q[(int64_t)c * M + y], e);
        Eref[(int64_t)y * n + i] = e;
        Ecur[(int64_t)y * n + i] = y0;
    }
}

__device__ __forceinline__ uint64_t wilcoxon_key(double eref, double ealt) {
    const double d = fabs(eref) - fabs(ealt);                           // pls.cpp:193
    const uint64_t mag = (uint64_t)__double_as_longlong(fabs(d));       // pls.cpp:198
    return (mag << 1) | (uint64_t)(d > 0.0);
}

// keys for tests (y, alt = a0 + b), b < B: segment index y*B + b. Advances Ecur by B components.
__global__ void keygen_kernel(const double* __restrict__ T, int64_t ldt, int64_t n, int M, int A, const double* __restrict__ Q,
                              const int* __restrict__ ref, const int* __restrict__ decided, int a0, int B,
                              const double* __restrict__ Eref, double* __restrict__ Ecur, uint64_t* __restrict__ keys,
                              int* __restrict__ seg_valid, long long* __restrict__ dsum) {
    const int y = blockIdx.y;
    const int ry = ref[y];
    const bool active = !decided[y] && a0 < ry;
    if (blockIdx.x == 0 && threadIdx.x < B) {
        seg_valid[y * B + threadIdx.x] = (active && a0 + (int)threadIdx.x < ry) ? 1 : 0;
        dsum[y * B + threadIdx.x] = 0;
    }
    if (!active) return;
    const int nb = min(B, ry - a0);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double er = Eref[(int64_t)y * n + i];
        double e = Ecur[(int64_t)y * n + i];
        for (int b = 0; b < nb; b++) {
            const int c = a0 + b;
            e = fma(-T[(int64_t)c * ldt + i], Q[(int64_t)c * M + y], e);
            keys[((int64_t)y * B + b) * n + i] = wilcoxon_key(er, e);
        }
        Ecur[(int64_t)y * n + i] = e;
    }
}

// d = sum_pos (pos+1) * sign  over a sorted segment (exact integer arithmetic; pls.cpp:202)
__global__ void __launch_bounds__(256) ranksum_kernel(const uint64_t* __restrict__ keys, int64_t n, const int* __restrict__ seg_valid,
                                                      long long* __restrict__ dsum) {
    const int seg = blockIdx.y;
    if (seg_valid && !seg_valid[seg]) return;
    __shared__ long long red[8];
    const uint64_t* k = keys + (int64_t)seg * n;
    long long acc = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t key = k[i];
        const long long s = (key == 0ull) ? 0ll : ((key & 1ull) ? 1ll : -1ll);
        acc += s * (long long)(i + 1);
    }
    acc = warp_sum_ll(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int w = 0; w < 8; w++) t += red[w];
        atomicAdd((unsigned long long*)&dsum[seg], (unsigned long long)t);
    }
}

// pls.cpp:152-160
__device__ __forceinline__ double normalcdf_dev(const double z) {
    const double c1 = 0.196854, c2 = 0.115194, c3 = 0.000344, c4 = 0.019527;
    const double zs = fabs(z);
    const double p = 0.5 / pow(1 + c1 * zs + c2 * zs * zs + c3 * zs * zs * zs + c4 * zs * zs * zs * zs, 4.0);
    return z < 0 ? p : 1.0 - p;
}
// pls.cpp:203-208 from the exact rank sum d
__device__ __forceinline__ double wilcoxon_p_from_d(long long d, unsigned long long n) {
    const double t = (double)(n * (n + 1)) / 2.0;
    const double v = (t - (double)d) / 2.0;
    const double ev = t / 2.0;
    const double sv = sqrt((double)(n * (n + 1) * (2 * n + 1)) / 24.0);
    const double z = (v - ev) / sv;
    return 1.0 - normalcdf_dev(z);
}

constexpr int SC_NB = 4096;
constexpr int SC_THREADS = 512;
constexpr int SC_BPT = SC_NB / SC_THREADS;   // buckets per thread in the scan

__device__ __forceinline__ long long rank_range_sum(long long a, long long b) {   // sum of ranks a..b inclusive (0 if empty)
    return (b >= a) ? (a + b) * (b - a + 1) / 2 : 0ll;
}

// status[seg]: 0 certain failure (p <= alpha), 1 certain success (p > alpha), 2 ambiguous (needs the exact sort)
__global__ void __launch_bounds__(SC_THREADS) wilcoxon_screen_kernel(const uint64_t* __restrict__ keys, int64_t n,
                                                                     const int* __restrict__ seg_valid, double alpha,
                                                                     int* __restrict__ status) {
    const int seg = blockIdx.x;
    if (!seg_valid[seg]) return;
    __shared__ uint32_t pos[SC_NB];
    __shared__ uint32_t neg[SC_NB];
    __shared__ double red[32];
    __shared__ long long lred[2][SC_THREADS / 32];
    __shared__ uint32_t wtot[SC_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint64_t* k = keys + (int64_t)seg * n;
    for (int i = tid; i < SC_NB; i += SC_THREADS) { pos[i] = 0; neg[i] = 0; }
    // pass 1: scale = mean |d| (deterministic block sum)
    double s = 0;
    for (int64_t i = tid; i < n; i += SC_THREADS) s += __longlong_as_double((long long)(k[i] >> 1));
    s = block_sum(s, red);
    const double mu = s / (double)n;
    // pass 2: signed bucket counts. bucket(x) = floor(NB * (1 - mu / (x + mu))): every step is monotone in x
    if (mu > 0.0) {   // mu == 0: every difference is zero, the histograms stay empty and d = 0 exactly
        for (int64_t i = tid; i < n; i += SC_THREADS) {
            const uint64_t key = k[i];
            if (key == 0ull) continue;   // zeros: lowest ranks, sign 0 (counted as n - sum of buckets)
            const double x = __longlong_as_double((long long)(key >> 1));
            int b = (int)((double)SC_NB * (1.0 - mu / (x + mu)));
            b = min(max(b, 0), SC_NB - 1);
            atomicAdd((key & 1ull) ? &pos[b] : &neg[b], 1u);
        }
    }
    __syncthreads();
    // exclusive scan of bucket populations: thread t owns buckets [t*SC_BPT, (t+1)*SC_BPT)
    uint32_t cnt = 0;
#pragma unroll
    for (int j = 0; j < SC_BPT; j++) cnt += pos[tid * SC_BPT + j] + neg[tid * SC_BPT + j];
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) wtot[wid] = incl;
    __syncthreads();
    uint32_t wbase = 0, total_nz = 0;
    for (int w = 0; w < SC_THREADS / 32; w++) { if (w < wid) wbase += wtot[w]; total_nz += wtot[w]; }
    // zeros hold the lowest ranks (key 0 is the smallest key) and carry sign 0
    long long R = (long long)((uint32_t)n - total_nz) + (long long)(wbase + incl - cnt);
    long long dlo = 0, dhi = 0;
#pragma unroll
    for (int j = 0; j < SC_BPT; j++) {
        const long long pb = pos[tid * SC_BPT + j], nb = neg[tid * SC_BPT + j], tb = pb + nb;
        if (tb) {
            dhi += rank_range_sum(R + tb - pb + 1, R + tb) - rank_range_sum(R + 1, R + nb);
            dlo += rank_range_sum(R + 1, R + pb) - rank_range_sum(R + tb - nb + 1, R + tb);
            R += tb;
        }
    }
    dlo = warp_sum_ll(dlo); dhi = warp_sum_ll(dhi);
    if (lane == 0) { lred[0][wid] = dlo; lred[1][wid] = dhi; }
    __syncthreads();
    if (tid == 0) {
        long long lo = 0, hi = 0;
        for (int w = 0; w < SC_THREADS / 32; w++) { lo += lred[0][w]; hi += lred[1][w]; }
        const double p_lo = wilcoxon_p_from_d(lo, (unsigned long long)n), p_hi = wilcoxon_p_from_d(hi, (unsigned long long)n);
        status[seg] = (p_lo > alpha) ? 1 : (!(p_hi > alpha) ? 0 : 2);
    }
}

// per y, walking the chunk in order: a certain success decides y; the first ambiguous test (and later ambiguous ones up
// to a certain success) are flagged for the exact path; exact_valid / dsum are prepared for it.
__global__ void screen_decide_kernel(const int* __restrict__ status, const int* __restrict__ seg_valid, int M, int B, int a0,
                                     int* __restrict__ decided, int* __restrict__ result, int* __restrict__ exact_valid,
                                     long long* __restrict__ dsum, int* __restrict__ n_exact) {
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= M) return;
    for (int b = 0; b < B; b++) { exact_valid[y * B + b] = 0; dsum[y * B + b] = 0; }
    if (decided[y]) return;
    bool pending = false;
    for (int b = 0; b < B; b++) {
        const int seg = y * B + b;
        if (!seg_valid[seg]) continue;
        const int st = status[seg];
        if (st == 1) { if (!pending) { decided[y] = 1; result[y] = a0 + b; } break; }
        if (st == 2) { pending = true; exact_valid[seg] = 1; atomicAdd(n_exact, 1); }
    }
}

// final decision of a chunk that needed exact tests: status 2 entries are replaced by the exact p-value
__global__ void decide_exact_kernel(const int* __restrict__ status, const long long* __restrict__ dsum, const int* __restrict__ seg_valid,
                                    const int* __restrict__ exact_valid, int M, int B, int a0, unsigned long long n, double alpha,
                                    int* __restrict__ decided, int* __restrict__ result) {
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= M || decided[y]) return;
    for (int b = 0; b < B; b++) {
        const int seg = y * B + b;
        if (!seg_valid[seg]) continue;
        int st = status[seg];
        if (st == 2) {
            if (!exact_valid[seg]) break;   // cannot happen: every ambiguous test before a success is flagged
            st = (wilcoxon_p_from_d(dsum[seg], n) > alpha) ? 1 : 0;
        }
        if (st == 1) { decided[y] = 1; result[y] = a0 + b; break; }
    }
}

__global__ void init_select_kernel(const int* __restrict__ ref, int M, int* __restrict__ decided, int* __restrict__ result) {
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= M) return;
    result[y] = ref[y];
    decided[y] = (ref[y] == 0) ? 1 : 0;
}

__global__ void single_keys_kernel(const double* __restrict__ e1, const double* __restrict__ e2, int64_t n, uint64_t* __restrict__ keys) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        keys[i] = wilcoxon_key(e1[i], e2[i]);
}
__global__ void single_p_kernel(const long long* __restrict__ dsum, unsigned long long n, double* __restrict__ p) {
    *p = wilcoxon_p_from_d(*dsum, n);
}

int press_grid(const abcb200_ctx* ctx, int64_t n) {
    const int64_t nblk = (n + PR_THREADS * PR_ROWS - 1) / (PR_THREADS * PR_ROWS);
    return (int)max((int64_t)1, min(nblk, (int64_t)(2 * ctx->sm_count)));
}

int select_bmax(int64_t n_te, int M) {
    const double per_b = (double)M * (double)n_te * 16.0;
    int b = (int)(1.5e9 / per_b);
    if (b < 1) b = 1;
    if (b > 32) b = 32;
    return b;
}

}  // namespace

size_t holdout_ws_bytes(const abcb200_ctx* ctx, int64_t n_te, int K, int M, int A) {
    size_t b = 0;
    const int64_t ldt = (n_te + 31) / 32 * 32;
    b += align_up((size_t)ldt * A * 8, 256);                                  // T
    b += align_up((size_t)press_grid(ctx, n_te) * M * A * 8, 256);           // PRESS partials
    b += align_up((size_t)M * A * 8, 256);                                    // PRESS
    b += 2 * align_up((size_t)n_te * M * 8, 256);                             // Eref, Ecur
    const int B = select_bmax(n_te, M);
    b += 2 * align_up((size_t)M * B * n_te * 8, 256);                         // keys, keys_alt
    b += radix_hist_bytes(n_te, M * B);
    b += 4 * align_up((size_t)M * 4, 256) + 2 * align_up((size_t)M * B * 8, 256) + 3 * align_up((size_t)M * B * 4, 256) + 512;
    return b + 8192;
}

int holdout_select_dev(abcb200_ctx* ctx, const double* Zte, int64_t ldx, const double* Yte, int64_t ldy, int64_t n_te,
                       const PlsFactors& f, double alpha, double* press_dev, int32_t* ncomp_host) {
    const int K = f.K, M = f.M, A = f.A;
    if (n_te <= 0) {   // empty hold-out: PRESS all zero -> argmin 0 -> one component for every response
        if (press_dev) CUDA_TRY(ctx, cudaMemsetAsync(press_dev, 0, sizeof(double) * M * A, ctx->stream));
        if (ncomp_host) for (int y = 0; y < M; y++) ncomp_host[y] = 1;
        return ABCB200_OK;
    }
    stage_begin(ctx, 2);
    const int64_t ldt = (n_te + 31) / 32 * 32;
    double* T = ws_new<double>(ctx, (size_t)ldt * A);
    const int pgrid = press_grid(ctx, n_te);
    double* partial = ws_new<double>(ctx, (size_t)pgrid * M * A);
    double* press = press_dev ? press_dev : ws_new<double>(ctx, (size_t)M * A);
    double* Eref = ws_new<double>(ctx, (size_t)n_te * M);
    double* Ecur = ws_new<double>(ctx, (size_t)n_te * M);
    const int Bmax = select_bmax(n_te, M);
    uint64_t* keys = ws_new<uint64_t>(ctx, (size_t)M * Bmax * n_te);
    uint64_t* keys_alt = ws_new<uint64_t>(ctx, (size_t)M * Bmax * n_te);
    uint32_t* hist = (uint32_t*)ws_alloc(ctx, radix_hist_bytes(n_te, M * Bmax));
    int* ref = ws_new<int>(ctx, M);
    int* decided = ws_new<int>(ctx, M);
    int* result = ws_new<int>(ctx, M);
    int* seg_valid = ws_new<int>(ctx, (size_t)M * Bmax);
    int* status = ws_new<int>(ctx, (size_t)M * Bmax);
    int* exact_valid = ws_new<int>(ctx, (size_t)M * Bmax);
    int* n_exact = ws_new<int>(ctx, 1);
    long long* dsum = ws_new<long long>(ctx, (size_t)M * Bmax);
    if (!T || !partial || !press || !Eref || !Ecur || !keys || !keys_alt || !hist || !ref || !decided || !result || !seg_valid || !dsum || !status || !exact_valid || !n_exact)
        ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in holdout_select");

    ABC_TRY(launch_xb(ctx, Zte, ldx, n_te, K, f.R, K, A, T, ldt));            // hold-out scores, all A components
    for (int y0 = 0; y0 < M; y0 += 32) {
        const int mc = min(32, M - y0);
        if (mc <= 8) LAUNCH(ctx, press_kernel<8>, pgrid, PR_THREADS, 0, T, ldt, Yte, ldy, n_te, M, A, f.Q, y0, partial);
        else if (mc <= 16) LAUNCH(ctx, press_kernel<16>, pgrid, PR_THREADS, 0, T, ldt, Yte, ldy, n_te, M, A, f.Q, y0, partial);
        else LAUNCH(ctx, press_kernel<32>, pgrid, PR_THREADS, 0, T, ldt, Yte, ldy, n_te, M, A, f.Q, y0, partial);
    }
    LAUNCH(ctx, press_reduce_kernel, (M * A + 255) / 256, 256, 0, partial, pgrid, M, A, 1.0, press);
    LAUNCH(ctx, argmin_kernel, (M + 3) / 4, 128, 0, press, M, A, ref);
    stage_end(ctx, 2);
    if (!ncomp_host) return ABCB200_OK;

    stage_begin(ctx, 3);
    LAUNCH(ctx, init_select_kernel, (M + 127) / 128, 128, 0, ref, M, decided, result);
    const int egrid = (int)max((int64_t)1, min((n_te + 255) / 256, (int64_t)(4 * ctx->sm_count)));
    LAUNCH(ctx, eref_kernel, dim3(egrid, M), 256, 0, T, ldt, Yte, ldy, n_te, M, f.Q, ref, Eref, Ecur);
    ABC_TRY(hpin_reserve(ctx, sizeof(int) * (3 * (size_t)M + 1) + 64));
    int* h_ref = (int*)ctx->hpin;
    int* h_decided = h_ref + M;
    int* h_result = h_decided + M;
    int* h_nexact = h_result + M;
    CUDA_TRY(ctx, cudaMemcpyAsync(h_ref, ref, sizeof(int) * M, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(h_decided, decided, sizeof(int) * M, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    int max_ref = 0;
    for (int y = 0; y < M; y++) max_ref = max(max_ref, h_ref[y]);
    int a0 = 0, B = min(4, Bmax);
    while (true) {
        bool any = false;
        for (int y = 0; y < M; y++) if (!h_decided[y] && h_ref[y] > a0) any = true;
        if (!any) break;
        LAUNCH(ctx, keygen_kernel, dim3(egrid, M), 256, 0, T, ldt, n_te, M, A, f.Q, ref, decided, a0, B, Eref, Ecur, keys, seg_valid, dsum);
        LAUNCH(ctx, wilcoxon_screen_kernel, M * B, SC_THREADS, 0, keys, n_te, seg_valid, alpha, status);
        CUDA_TRY(ctx, cudaMemsetAsync(n_exact, 0, sizeof(int), ctx->stream));
        LAUNCH(ctx, screen_decide_kernel, (M + 127) / 128, 128, 0, status, seg_valid, M, B, a0, decided, result, exact_valid, dsum, n_exact);
        CUDA_TRY(ctx, cudaMemcpyAsync(h_decided, decided, sizeof(int) * M, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(h_nexact, n_exact, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (*h_nexact > 0) {   // some intervals straddle the threshold: sort exactly those tests
            ABC_TRY(radix_sort_segments(ctx, keys, keys_alt, nullptr, nullptr, n_te, M * B, hist, exact_valid));
            const int rgrid = (int)max((int64_t)1, min((n_te + 2047) / 2048, (int64_t)64));
            LAUNCH(ctx, ranksum_kernel, dim3(rgrid, M * B), 256, 0, keys, n_te, exact_valid, dsum);
            LAUNCH(ctx, decide_exact_kernel, (M + 127) / 128, 128, 0, status, dsum, seg_valid, exact_valid, M, B, a0, (unsigned long long)n_te, alpha, decided, result);
            CUDA_TRY(ctx, cudaMemcpyAsync(h_decided, decided, sizeof(int) * M, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        }
        a0 += B;
        B = min(2 * B, Bmax);
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(h_result, result, sizeof(int) * M, cudaMemcpyDeviceToHost, ctx->stream));
    stage_end(ctx, 3);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (int y = 0; y < M; y++) ncomp_host[y] = h_result[y] + 1;   // index -> component count (pls.cpp:288)
    return ABCB200_OK;
}

size_t wilcoxon_ws_bytes(int64_t n) { return 2 * align_up((size_t)n * 8, 256) + radix_hist_bytes(n, 1) + 1024; }

int wilcoxon_dev(abcb200_ctx* ctx, const double* e1, const double* e2, int64_t n, double* p_host) {
    uint64_t* keys = ws_new<uint64_t>(ctx, n);
    uint64_t* keys_alt = ws_new<uint64_t>(ctx, n);
    uint32_t* hist = (uint32_t*)ws_alloc(ctx, radix_hist_bytes(n, 1));
    long long* dsum = ws_new<long long>(ctx, 1);
    double* p = ws_new<double>(ctx, 1);
    if (!keys || !keys_alt || !hist || !dsum || !p) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in wilcoxon");
    CUDA_TRY(ctx, cudaMemsetAsync(dsum, 0, sizeof(long long), ctx->stream));
    const int grid = (int)max((int64_t)1, min((n + 255) / 256, (int64_t)(4 * ctx->sm_count)));
    LAUNCH(ctx, single_keys_kernel, grid, 256, 0, e1, e2, n, keys);
    ABC_TRY(radix_sort_segments(ctx, keys, keys_alt, nullptr, nullptr, n, 1, hist, nullptr));
    LAUNCH(ctx, ranksum_kernel, dim3((int)max((int64_t)1, min((n + 2047) / 2048, (int64_t)64)), 1), 256, 0, keys, n, (const int*)nullptr, dsum);
    LAUNCH(ctx, single_p_kernel, 1, 1, 0, dsum, (unsigned long long)n, p);
    ABC_TRY(hpin_reserve(ctx, 64));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hpin, p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *p_host = *(double*)ctx->hpin;
    return ABCB200_OK;
}
