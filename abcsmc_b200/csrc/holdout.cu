// holdout.cu — S3 hold-out residuals + PRESS and S4 Wilcoxon component selection.
//
// Reference: Model::cv_NEW_DATA lib/PLS/src/pls.cpp:494-510, PLS::validation :235-261,
//            PLS::optimal_num_components :265-289, PLS::wilcoxon :190-211, PLS::normalcdf :152-160.
//
// The reference materialises an M x n_te x A error cube (4.5 GB at the dengue config) from A dense products
// Y - X (R_c Q_c^T). Here the hold-out scores T = X_te R are formed once (DMMA, launch_xb) and residuals are the
// running prefix e_c = e_{c-1} - t_c q_c^T. One pass (press_chk_kernel) accumulates PRESS for every (y, c) and stores
// the residual after every G-th component (checkpoints), so any error column is at most G FMAs away from HBM.
//
// Selection (pls.cpp:276-287): per response y, ref = first argmin PRESS, then the SMALLEST alt < ref with
// wilcoxon(E[:,ref], E[:,alt]) > alpha. Only the decision p > alpha is consumed and p is monotone in the signed rank
// sum d, so every test (y, alt < ref) is first BRACKETED instead of sorted: the |differences| are binned by a monotone
// map, which pins every element's rank to its bin and d to an exact integer interval [d_lo, d_hi] (positives at the
// bottom / top of each bin). p(d_lo) > alpha: certain success; p(d_hi) <= alpha: certain failure.
//   level 1 (screen1_kernel): 64 bins, per-thread private u8 counters in shared memory (plain LDS/STS, no atomics), four
//                             tests per row pass; decides every test whose |z| is more than ~13 away from the threshold;
//   level 2 (screen2_kernel): 4096 bins of (nearly) equal mass derived from the level-1 counts, shared-memory atomics,
//                             one CTA per remaining test (interval width ~0.1 sigma);
//   level 3: the tests still straddling the threshold are sorted exactly (segmented radix sort, sort.cu) and
//            d = sum rank * sign is evaluated in integer arithmetic.
// Decisions are therefore identical to sorting every test. Keys: bits(|d|) << 1 | (d > 0), d = |e_ref| - |e_alt|;
// d == 0 gives key 0 and sign 0 (pls.cpp:196). Exact ties in |d| between opposite signs are ordered negative-first
// (the reference's order there is whatever introsort yields).
// Everything up to the end of level 2 is enqueued without a host round trip; one D2H of (results, #exact) follows.
#include <stdlib.h>

#include "kernels.cuh"

namespace {

constexpr int CHK_G = 4;          // a checkpoint after every CHK_G components
constexpr int PC_THREADS = 128;
constexpr int PC_MY = 8;          // responses per CTA (register tile)

// PRESS partials + checkpoints. 1-D grid of nblk * ycta CTAs, CTA = (row block of PC_ROWS rows, PC_MY responses); the
// response blocks of the same rows are neighbours in the grid, so their reads of T meet in L2. A thread owns two rows for
// the whole kernel and carries their residuals e_c = e_{c-1} - t_c q_c (pls.cpp:449-455 as a running prefix) in registers
// through ALL A components: Y is read once, T once per response block, and nothing is read back. The scores of the next
// CHK_G components are requested one iteration ahead. PRESS: per iteration the warp's 32 (response, component) sums come
// out of one transposing butterfly into a per-warp shared-memory slot; every PC_FLUSH iterations the CTA adds the warps
// up in a fixed order (deterministic) and writes the block's partial sums.
// chk[((y * (nchk - 1)) + k - 1) * ldn + i] = residual of response y, row i after k*CHK_G components (k = 1 .. nchk-1).
// partial[blk * M * A + y * A + c] = sum over the block's rows of e_c[i, y]^2.
constexpr int PC_ROWS = 2 * PC_THREADS;
constexpr int PC_FLUSH = 32;
constexpr int PC_PF = 3;           // L2 prefetch distance of the scores, in iterations of CHK_G components
constexpr size_t PC_SMEM = sizeof(double) * ((size_t)PC_FLUSH * CHK_G * PC_MY + (size_t)(PC_THREADS / 32) * PC_FLUSH * 32);
__global__ void __launch_bounds__(PC_THREADS, 3) press_chk_kernel(const double* __restrict__ T, int64_t ldt, const double* __restrict__ Y,
                                                                  int64_t ldy, int64_t n, int M, int A, const double* __restrict__ Q,
                                                                  int ycta, int nchk, int64_t ldn, double* __restrict__ chk,
                                                                  double* __restrict__ partial, int k_begin, int k_end) {
    extern __shared__ __align__(16) double pc_sm[];
    double (*qs)[PC_MY] = (double (*)[PC_MY])pc_sm;                                        // [PC_FLUSH * CHK_G][PC_MY]
    double* wred = pc_sm + PC_FLUSH * CHK_G * PC_MY;                                       // [warp][PC_FLUSH][32]
    static_assert(PC_MY * CHK_G == 32, "one lane per accumulator");
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t blk = blockIdx.x / ycta;
    const int y0 = (int)(blockIdx.x % ycta) * PC_MY;
    const int my = min(PC_MY, M - y0);
    const int64_t i1 = blk * PC_ROWS + tid, i2 = i1 + PC_THREADS;
    const bool v1 = i1 < n, v2 = i2 < n;
    double* out = partial + blk * M * A;
    // rows past the end carry e = 0 and t = 0: they add exact zeros to every sum
    // Checkpoint units [k_begin, k_end) of this launch (the pipelined ranking feeds the kernel blocks of components as the fit
    // emits them): the residuals start from Y, or from the checkpoint the previous launch stored after k_begin * CHK_G components.
    const int64_t chk_ystride = (int64_t)(nchk - 1) * ldn;
    double e[PC_MY], f[PC_MY];
#pragma unroll
    for (int yy = 0; yy < PC_MY; yy++) {
        const double* yp = (k_begin == 0) ? Y + (int64_t)(y0 + min(yy, my - 1)) * ldy
                                          : chk + ((int64_t)(y0 + min(yy, my - 1)) * (nchk - 1) + (k_begin - 1)) * ldn;
        e[yy] = (v1 && yy < my) ? yp[i1] : 0.0;
        f[yy] = (v2 && yy < my) ? yp[i2] : 0.0;
    }
    double tn[CHK_G], un[CHK_G];
    auto load_scores = [&](int c0) {
#pragma unroll
        for (int cc = 0; cc < CHK_G; cc++) {
            const bool cv = c0 + cc < A;
            const double* tp = T + (int64_t)min(c0 + cc, A - 1) * ldt;
            tn[cc] = (cv && v1) ? tp[i1] : 0.0;
            un[cc] = (cv && v2) ? tp[i2] : 0.0;
        }
    };
    load_scores(k_begin * CHK_G);
    for (int kb = k_begin; kb < k_end; kb += PC_FLUSH) {
        const int kn = min(PC_FLUSH, k_end - kb);
        for (int idx = tid; idx < kn * CHK_G * PC_MY; idx += PC_THREADS) {
            const int cc = idx / PC_MY, yy = idx % PC_MY, c = kb * CHK_G + cc;
            qs[cc][yy] = (c < A && yy < my) ? Q[(int64_t)c * M + y0 + yy] : 0.0;
        }
        __syncthreads();
        for (int kk = 0; kk < kn; kk++) {
            const int k = kb + kk;
            double t[CHK_G], u[CHK_G];
#pragma unroll
            for (int cc = 0; cc < CHK_G; cc++) { t[cc] = tn[cc]; u[cc] = un[cc]; }
            if (k + 1 < k_end) load_scores((k + 1) * CHK_G);
            if (k + PC_PF < k_end && (lane & 15) == 0) {       // the score lines PC_PF iterations ahead, into L2 (one lane per 128-byte line)
#pragma unroll
                for (int cc = 0; cc < CHK_G; cc++) {
                    const double* tp = T + (int64_t)min((k + PC_PF) * CHK_G + cc, A - 1) * ldt;
                    if (v1) asm volatile("prefetch.global.L2 [%0];" ::"l"(tp + i1));
                    if (v2) asm volatile("prefetch.global.L2 [%0];" ::"l"(tp + i2));
                }
            }
            double acc[32];
#pragma unroll
            for (int cc = 0; cc < CHK_G; cc++) {
#pragma unroll
                for (int yy = 0; yy < PC_MY; yy++) {
                    const double q = qs[kk * CHK_G + cc][yy];
                    e[yy] = fma(-t[cc], q, e[yy]);
                    f[yy] = fma(-u[cc], q, f[yy]);
                    acc[yy * CHK_G + cc] = fma(f[yy], f[yy], e[yy] * e[yy]);
                }
            }
            if (k + 1 < nchk) {      // the residual after (k + 1) * CHK_G components is checkpoint k + 1
                double* sp = chk + ((int64_t)y0 * (nchk - 1) + k) * ldn;
#pragma unroll
                for (int yy = 0; yy < PC_MY; yy++) if (yy < my) {
                    if (v1) sp[yy * chk_ystride + i1] = e[yy];
                    if (v2) sp[yy * chk_ystride + i2] = f[yy];
                }
            }
            wred[((size_t)wid * PC_FLUSH + kk) * 32 + lane] = warp_sum32_transposed(acc);
        }
        __syncthreads();
        for (int idx = tid; idx < kn * 32; idx += PC_THREADS) {
            const int kk = idx >> 5, l = idx & 31, yy = l / CHK_G, c = (kb + kk) * CHK_G + (l % CHK_G);
            double sacc = 0.0;
#pragma unroll
            for (int w = 0; w < PC_THREADS / 32; w++) sacc += wred[((size_t)w * PC_FLUSH + kk) * 32 + l];
            if (yy < my && c < A) out[(int64_t)(y0 + yy) * A + c] = sacc;
        }
        __syncthreads();
    }
}

// press[y, c] = sum over row blocks (fixed order). CTA = (response y, 32 components): 32 component lanes x 16 block slices, every
// thread sums its slice with four loads in flight, the slices are combined in a fixed order (deterministic). One CTA per
// response over all A components took 0.45 ms at 1M particles (1953 row blocks walked by 2 slices).
constexpr int PF_T = 512;
constexpr int PF_CW = 32;
__global__ void __launch_bounds__(PF_T) press_reduce_kernel(const double* __restrict__ partial, int nblk, int M, int A, double* __restrict__ press) {
    __shared__ double part[PF_T];
    const int y = blockIdx.x, tid = threadIdx.x;
    constexpr int BS = PF_T / PF_CW;
    const int cs = tid % PF_CW, bs = tid / PF_CW, c = blockIdx.y * PF_CW + cs;
    double a[4] = {0, 0, 0, 0};                    // four interleaved partial sums (fixed order), loads in flight
    if (c < A) {
        const double* pp = partial + (int64_t)y * A + c;
        int b = bs;
        for (; b + 3 * BS < nblk; b += 4 * BS) {
#pragma unroll
            for (int u = 0; u < 4; u++) a[u] += pp[(int64_t)(b + u * BS) * M * A];
        }
#pragma unroll
        for (int u = 0; u < 4; u++) if (b + u * BS < nblk) a[u] += pp[(int64_t)(b + u * BS) * M * A];
    }
    part[tid] = (a[0] + a[1]) + (a[2] + a[3]);
    __syncthreads();
    if (bs == 0 && c < A) {
        double sacc = 0.0;
        for (int j = 0; j < BS; j++) sacc += part[j * PF_CW + cs];
        press[(int64_t)c * M + y] = sacc;
    }
}

// ref[y] = first argmin_c press[y, c] (Eigen minCoeff(&idx), pls.cpp:278); selection state initialised (result = ref, decided iff
// ref == 0). One warp per response.
__global__ void __launch_bounds__(128) press_argmin_kernel(const double* __restrict__ press, int M, int A, int* __restrict__ ref, int* __restrict__ decided,
                                                           int* __restrict__ result) {
    const int y = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (y >= M) return;
    double best = 0; int besti = 0x7fffffff;
    for (int c = lane; c < A; c += 32) {           // c ascending per lane: keeps the lane's first minimum
        const double v = press[(int64_t)c * M + y];
        if (besti == 0x7fffffff || v < best) { best = v; besti = c; }
    }
    // first arg-min over the lanes: smaller value, then smaller index (lanes without a component carry index INT_MAX)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (oi != 0x7fffffff && (besti == 0x7fffffff || ov < best || (ov == best && oi < besti))) { best = ov; besti = oi; }
    }
    if (lane == 0) { ref[y] = besti; result[y] = besti; decided[y] = (besti == 0) ? 1 : 0; }
}

// residual of response y, row i after `ncomp` components, from the nearest checkpoint below
__device__ __forceinline__ double residual_at(const double* __restrict__ T, int64_t ldt, const double* __restrict__ Y, int64_t ldy,
                                              const double* __restrict__ chk, int nchk, int64_t ldn, const double* __restrict__ Q, int M,
                                              int y, int64_t i, int ncomp) {
    int k = ncomp / CHK_G;
    if (k > nchk - 1) k = nchk - 1;
    double e = (k == 0) ? Y[(int64_t)y * ldy + i] : chk[((int64_t)y * (nchk - 1) + k - 1) * ldn + i];
    for (int c = k * CHK_G; c < ncomp; c++) e = fma(-T[(int64_t)c * ldt + i], Q[(int64_t)c * M + y], e);
    return e;
}

// Eref[y * ldn + i] = residual with ref[y] + 1 components
__global__ void eref_kernel(const double* __restrict__ T, int64_t ldt, const double* __restrict__ Y, int64_t ldy, int64_t n, int M,
                            const double* __restrict__ chk, int nchk, int64_t ldn, const double* __restrict__ Q,
                            const int* __restrict__ ref, double* __restrict__ Eref) {
    const int y = blockIdx.y;
    const int nc = ref[y] + 1;
    if (nc <= 1) return;                      // ref == 0: no test will be run for this response
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        Eref[(int64_t)y * ldn + i] = residual_at(T, ldt, Y, ldy, chk, nchk, ldn, Q, M, y, i, nc);
}

__device__ __forceinline__ uint64_t wilcoxon_key(double eref, double ealt) {
    const double d = fabs(eref) - fabs(ealt);                           // pls.cpp:193
    const uint64_t mag = (uint64_t)__double_as_longlong(fabs(d));       // pls.cpp:198
    return (mag << 1) | (uint64_t)(d > 0.0);
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
__device__ __forceinline__ long long rank_range_sum(long long a, long long b) {   // sum of ranks a..b inclusive (0 if empty)
    return (b >= a) ? (a + b) * (b - a + 1) / 2 : 0ll;
}
// contribution of one bin (pb positives, nb negatives, first rank R + 1) to the lower / upper bound of d
__device__ __forceinline__ void bin_bounds(long long R, long long pb, long long nb, long long& dlo, long long& dhi) {
    const long long tb = pb + nb;
    dhi += rank_range_sum(R + tb - pb + 1, R + tb) - rank_range_sum(R + 1, R + nb);
    dlo += rank_range_sum(R + 1, R + pb) - rank_range_sum(R + tb - nb + 1, R + tb);
}
__device__ __forceinline__ int status_from_bounds(long long dlo, long long dhi, unsigned long long n, double alpha) {
    const double p_lo = wilcoxon_p_from_d(dlo, n), p_hi = wilcoxon_p_from_d(dhi, n);
    return (p_lo > alpha) ? 1 : (!(p_hi > alpha) ? 0 : 2);   // 1 certain success, 0 certain failure, 2 ambiguous
}

// ---- level 1 ---------------------------------------------------------------------------------------------------------
// One CTA = the four tests alt = 4g .. 4g+3 of one response over a range of rows. A thread owns whole rows: it loads the
// checkpoint after 4g components, the reference residual and four T values, and gets all four error columns with four
// FMAs (1.5 loads per test and row). Its 4 x 64 x 2 bin counters are private u8 cells in shared memory laid out so that
// lane l only ever touches bank l (plain LDS/STS, no atomics); they are folded into u32 totals every 255 rows.
// Row splits of the same group merge their totals with global atomics; the last CTA to arrive evaluates the brackets.
constexpr int S1_THREADS = 128;                     // 64 KB of counters per CTA: three CTAs per SM, out of phase with each other
constexpr int S1_TESTS = CHK_G;                    // tests per CTA = components per checkpoint interval
constexpr int S1_NB = 64;                          // bins
constexpr int S1_WORDS = S1_TESTS * S1_NB * 2 / 4; // 32-bit words of counters per thread (4 u8 cells per word)
constexpr int S1_ROWS = 4;                         // rows per thread and trip (loads in flight: 6 per row)
constexpr int S1_SAMPLE_ROWS = 1024;               // rows sampled for the bin scale
constexpr size_t S1_SMEM = (size_t)S1_WORDS * S1_THREADS * 4;   // 64 KB
constexpr int S1_CTAS_PER_SM = 3;
#ifndef S1_PF_TRIPS
#define S1_PF_TRIPS 3
#endif
constexpr int S1_PF = S1_PF_TRIPS;                  // L2 prefetch distance in trips (see prefetch_trip)
static_assert(S1_TESTS == 4 && S1_WORDS == 128, "counter layout assumes 4 tests x 64 bins x 2 signs");
static_assert(S1_THREADS >= S1_WORDS && S1_THREADS >= 32 * S1_TESTS, "fold() uses one thread per counter word, the brackets one warp per test");

struct TestInfo {          // per test (y * A + alt), written by level 1 for the tests it leaves ambiguous
    double scale;          // bin = min((int)(|d| * scale), S1_NB - 1)
    unsigned int zeros;    // number of exactly-zero differences (lowest ranks, sign 0)
    unsigned int pad;
    unsigned int pos[S1_NB], neg[S1_NB];
};

// counter cell of (test s, bin b, sign sg) for thread t: word (s * 32 + b / 2) of the thread's column (words of one index are S1_THREADS
// apart, so a warp's accesses hit 32 different banks), byte (b & 1) * 2 + sg of that word

// ghist: [group][test][sign][bin] u32 totals (+ [4] zero counts) shared by the row splits of a group; ticket: arrivals
constexpr int S1_GH = S1_TESTS * 2 * S1_NB + S1_TESTS;

__global__ void __launch_bounds__(S1_THREADS, S1_CTAS_PER_SM) screen1_kernel(const double* __restrict__ T, int64_t ldt, const double* __restrict__ Y,
                                                                int64_t ldy, int64_t n, int M, int A, const double* __restrict__ chk,
                                                                int nchk, int64_t ldn, const double* __restrict__ Q,
                                                                const double* __restrict__ Eref, const int* __restrict__ ref,
                                                                double alpha, int64_t rows_per_split, unsigned int* __restrict__ ghist,
                                                                unsigned int* __restrict__ ticket, int* __restrict__ status,
                                                                TestInfo* __restrict__ info, int pf_trips) {
    extern __shared__ __align__(16) unsigned char cells[];
    __shared__ unsigned int tot[S1_GH];
    __shared__ double sred[S1_TESTS][S1_THREADS / 32];
    __shared__ double s_scale[S1_TESTS];
    __shared__ int s_last;
    // grid = (responses, groups, row splits): the M CTAs that share a group's four score columns are neighbours in launch order and
    // run side by side, so the columns are read from HBM once and from L2 by the others (with the groups of ONE response as
    // neighbours the resident CTAs touched all of T plus a dozen responses' checkpoints, ~4x L2, and T was evicted between uses)
    const int y = blockIdx.x, g = blockIdx.y;
    const int ry = ref[y];
    const int a0 = g * S1_TESTS;
    if (a0 >= ry) return;                                      // whole CTA: uniform
    const int ntest = min(S1_TESTS, ry - a0);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int ngroup = gridDim.y;
    const double* e0p = (g == 0) ? Y + (int64_t)y * ldy : chk + ((int64_t)y * (nchk - 1) + g - 1) * ldn;
    const double* erp = Eref + (int64_t)y * ldn;
    const double* tp[S1_TESTS];
    double qy[S1_TESTS];
#pragma unroll
    for (int s = 0; s < S1_TESTS; s++) { const int c = min(a0 + s, A - 1); tp[s] = T + (int64_t)c * ldt; qy[s] = (s < ntest) ? Q[(int64_t)c * M + y] : 0.0; }
    for (int i = tid; i < S1_GH; i += S1_THREADS) tot[i] = 0;
    unsigned int* mywords = (unsigned int*)cells;
    for (int wd = 0; wd < S1_WORDS; wd++) mywords[(size_t)wd * S1_THREADS + tid] = 0u;

    // ---- bin scale per test from the first rows of the set (identical in every row split) -----------------------------
    double sm[S1_TESTS] = {0, 0, 0, 0};
    const int64_t nsamp = min((int64_t)S1_SAMPLE_ROWS, n);
    for (int64_t i = tid; i < nsamp; i += S1_THREADS) {
        double e = e0p[i];
        const double er = fabs(erp[i]);
#pragma unroll
        for (int s = 0; s < S1_TESTS; s++) { e = fma(-tp[s][i], qy[s], e); sm[s] += fabs(er - fabs(e)); }
    }
#pragma unroll
    for (int s = 0; s < S1_TESTS; s++) { const double v = warp_sum(sm[s]); if (lane == 0) sred[s][wid] = v; }
    __syncthreads();
    if (tid < S1_TESTS) {
        double mu = 0;
        for (int w = 0; w < S1_THREADS / 32; w++) mu += sred[tid][w];
        mu /= (double)nsamp;
        s_scale[tid] = (mu > 0.0 && mu < 1e300) ? (double)S1_NB / (6.0 * mu) : 0.0;
    }
    __syncthreads();
    double scale[S1_TESTS];
#pragma unroll
    for (int s = 0; s < S1_TESTS; s++) scale[s] = s_scale[s];

    // ---- stream the rows of this split ---------------------------------------------------------------------------------
    const int64_t r0 = (int64_t)blockIdx.z * rows_per_split, r1 = min(n, r0 + rows_per_split);
    unsigned int zeros[S1_TESTS] = {0, 0, 0, 0};
    auto fold = [&]() {   // add every thread's u8 cells to the u32 totals and clear them; word wd -> (test, 2 bins, 2 signs)
        __syncthreads();
        if (tid < S1_WORDS) {
            unsigned int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
            for (int j = 0; j < S1_THREADS; j++) {
                const unsigned int v = mywords[(size_t)tid * S1_THREADS + ((j + tid) & (S1_THREADS - 1))];   // rotated: one bank per lane
                c0 += v & 0xffu; c1 += (v >> 8) & 0xffu; c2 += (v >> 16) & 0xffu; c3 += v >> 24;
            }
            const int s = tid >> 5, b = (tid & 31) * 2;
            unsigned int* tt = tot + s * 2 * S1_NB;            // [sign][bin]
            tt[b] += c0; tt[S1_NB + b] += c1; tt[b + 1] += c2; tt[S1_NB + b + 1] += c3;
        }
        __syncthreads();
        for (int wd = 0; wd < S1_WORDS; wd++) mywords[(size_t)wd * S1_THREADS + tid] = 0u;
    };
    int since_fold = 0;
    // Software pipeline: the loads of trip n+1 are issued before trip n is binned, so HBM requests stay in flight while the
    // dependent LDS -> add -> STS chains of the private counters run (16 per trip).
    double en[S1_ROWS], ern[S1_ROWS], tvn[S1_ROWS][S1_TESTS];
    auto load_trip = [&](int64_t b0) {
#pragma unroll
        for (int r = 0; r < S1_ROWS; r++) {
            const int64_t ic = b0 + tid + (int64_t)r * S1_THREADS;
            en[r] = e0p[ic]; ern[r] = erp[ic];
#pragma unroll
            for (int s = 0; s < S1_TESTS; s++) tvn[r][s] = tp[s][ic];
        }
    };
    // The register prefetch above reaches one trip (4 rows) ahead and the warps still wait on it (long scoreboard, half of all samples:
    // profiles/r02_lines_C3_screen1_kernel_n1.txt); a deeper register pipeline does not fit (157 of 168 registers, shared memory full
    // of counters). So the lines of the trip after the next one are pulled into L2 on the side: one lane per 128-byte line.
    auto prefetch_trip = [&](int64_t b0) {
        if ((lane & 15) == 0) {
#pragma unroll
            for (int r = 0; r < S1_ROWS; r++) {
                const int64_t ic = b0 + tid + (int64_t)r * S1_THREADS;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(e0p + ic));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(erp + ic));
#pragma unroll
                for (int s = 0; s < S1_TESTS; s++) asm volatile("prefetch.global.L2 [%0];" ::"l"(tp[s] + ic));
            }
        }
    };
    // Lean inner step (the kernel is bound by instruction issue, not by HBM or the LDS/STS rate): per (row, test) one DFMA (the
    // next error), one DADD (the difference, pls.cpp:193), one DMUL + saturating F2I + integer min (the bin; the same map as
    // `(int)fmin(|d| * scale, 63)`), the sign bit of d, four integer instructions for the cell address, LDS.U8 / +1 / STS.U8.
    // Exact zeros (sign 0, lowest ranks) land in cell (bin 0, +) like +0.0 and are counted on the side; the evaluation takes them
    // out of that cell again. Tests past `ntest` (last group of a response) count into cells nobody reads.
    unsigned char* const mycells = cells + (size_t)tid * 4;
    auto bin_row = [&](const double e0, const double er0, const double (&tvr)[S1_TESTS]) {
        const double aer = fabs(er0);
        double ee = e0;
        unsigned char* cp[S1_TESTS];
#pragma unroll
        for (int s = 0; s < S1_TESTS; s++) {
            ee = fma(-tvr[s], qy[s], ee);
            const double d = aer - fabs(ee);               // pls.cpp:193
            zeros[s] += (d == 0.0) ? 1u : 0u;
            const unsigned int b = (unsigned int)min(__double2int_rz(fabs(d) * scale[s]), S1_NB - 1);
            const unsigned int c = 2u * b + ((unsigned int)__double2hiint(d) >> 31);     // cell index: (bin, sign)
            cp[s] = mycells + s * (S1_WORDS / S1_TESTS) * S1_THREADS * 4 + ((c & ~3u) << 7) + (c & 3u);   // word (c >> 2) of test s, byte c & 3
        }
        // the four cells of a row belong to four different tests (disjoint words): read together, written together
        unsigned int cv[S1_TESTS];
#pragma unroll
        for (int s = 0; s < S1_TESTS; s++) cv[s] = *cp[s];
#pragma unroll
        for (int s = 0; s < S1_TESTS; s++) *cp[s] = (unsigned char)(cv[s] + 1u);
    };
    constexpr int64_t S1_TRIP = (int64_t)S1_ROWS * S1_THREADS;
    const int64_t full_end = r0 + ((r1 > r0) ? (r1 - r0) / S1_TRIP * S1_TRIP : 0);      // full trips: no bounds checks
    if (r0 < full_end) load_trip(r0);
    for (int64_t b0 = r0; b0 < full_end; b0 += S1_TRIP) {                               // uniform trip count (fold() has barriers)
        double e[S1_ROWS], er[S1_ROWS], tv[S1_ROWS][S1_TESTS];
#pragma unroll
        for (int r = 0; r < S1_ROWS; r++) {
            e[r] = en[r]; er[r] = ern[r];
#pragma unroll
            for (int s = 0; s < S1_TESTS; s++) tv[r][s] = tvn[r][s];
        }
        if (b0 + S1_TRIP < full_end) load_trip(b0 + S1_TRIP);
        if (b0 + pf_trips * S1_TRIP < full_end) prefetch_trip(b0 + pf_trips * S1_TRIP);
#pragma unroll
        for (int r = 0; r < S1_ROWS; r++) bin_row(e[r], er[r], tv[r]);
        since_fold += S1_ROWS;
        if (since_fold + S1_ROWS > 255) { fold(); since_fold = 0; }   // trip counts are uniform across the CTA
    }
    for (int64_t i = full_end + tid; i < r1; i += S1_THREADS) {      // the ragged tail of the split: < S1_ROWS rows per thread
        double tvr[S1_TESTS];
#pragma unroll
        for (int s = 0; s < S1_TESTS; s++) tvr[s] = tp[s][i];
        bin_row(e0p[i], erp[i], tvr);
    }
    fold();
#pragma unroll
    for (int s = 0; s < S1_TESTS; s++) { const unsigned int z = (unsigned int)warp_sum_ll((long long)zeros[s]); if (lane == 0 && z) atomicAdd(&tot[S1_TESTS * 2 * S1_NB + s], z); }
    __syncthreads();

    // ---- merge the row splits; the last CTA of the group evaluates the brackets ------------------------------------------
    const unsigned int* h = tot;
    if (gridDim.z > 1) {
        unsigned int* gh = ghist + ((size_t)y * ngroup + g) * S1_GH;
        for (int i = tid; i < S1_GH; i += S1_THREADS) if (tot[i]) atomicAdd(&gh[i], tot[i]);
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(&ticket[(size_t)y * ngroup + g], 1u) == gridDim.z - 1) ? 1 : 0;
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        for (int i = tid; i < S1_GH; i += S1_THREADS) tot[i] = __ldcg(&gh[i]);
        __syncthreads();
    }
    if (wid < ntest) {   // warp s: rank brackets of test a0 + s from its 64 bins (2 per lane)
        const int s = wid;
        const unsigned int* hp = h + s * 2 * S1_NB;
        const unsigned int nz = h[S1_TESTS * 2 * S1_NB + s];
        const long long p0 = (long long)hp[2 * lane] - ((lane == 0) ? (long long)nz : 0ll);   // exact zeros were binned as (bin 0, +)
        const long long n0 = hp[S1_NB + 2 * lane], p1 = hp[2 * lane + 1], n1 = hp[S1_NB + 2 * lane + 1];
        const unsigned int mine = (unsigned int)(p0 + n0 + p1 + n1);
        unsigned int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        const long long R = (long long)nz + (long long)(incl - mine);
        long long dlo = 0, dhi = 0;
        bin_bounds(R, p0, n0, dlo, dhi);
        bin_bounds(R + p0 + n0, p1, n1, dlo, dhi);
        dlo = warp_sum_ll(dlo); dhi = warp_sum_ll(dhi);
        const int st = status_from_bounds(dlo, dhi, (unsigned long long)n, alpha);
        const int64_t test = (int64_t)y * A + a0 + s;
        if (lane == 0) status[test] = st;
        if (st == 2) {
            TestInfo* ti = info + test;
            if (lane == 0) { ti->scale = s_scale[s]; ti->zeros = nz; ti->pad = 0; }
            ti->pos[2 * lane] = (unsigned int)p0; ti->pos[2 * lane + 1] = (unsigned int)p1;
            ti->neg[2 * lane] = (unsigned int)n0; ti->neg[2 * lane + 1] = (unsigned int)n1;
        }
    }
}

// Per response, walking alt ascending (pls.cpp:281-286): a certain success with no ambiguous test before it decides y.
// Ambiguous tests met before that are appended to the work list of the next level. work[0] = count, entries from work[1].
// One warp per response, 32 statuses per step (a thread per response walked up to A dependent loads: 42 us per launch at A = 150).
// The work list keeps alt ascending within a response (a warp reserves its slots with one atomic per step).
__global__ void __launch_bounds__(1024) decide_kernel(const int* __restrict__ status, const int* __restrict__ ref, int M, int A, int* __restrict__ decided,
                                                     int* __restrict__ result, int* __restrict__ work, int* __restrict__ summ, const int* __restrict__ other_work) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int y = blockIdx.x * nw + wid; y < M; y += gridDim.x * nw) {
        if (decided[y]) continue;
        const int ry = ref[y];
        bool pending = false, done = false;
        for (int a0 = 0; a0 < ry && !done; a0 += 32) {
            const int alt = a0 + lane;
            const int st = (alt < ry) ? status[(int64_t)y * A + alt] : 0;
            const unsigned succ = __ballot_sync(0xffffffffu, st == 1);
            const int first_succ = succ ? __ffs(succ) - 1 : 32;                  // lanes at or past it are not looked at
            const unsigned amb = __ballot_sync(0xffffffffu, st == 2 && lane < first_succ);
            if (amb) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&work[0], __popc(amb));
                base = __shfl_sync(0xffffffffu, base, 0);
                if ((amb >> lane) & 1u) work[1 + base + __popc(amb & ((1u << lane) - 1u))] = y * A + alt;
                pending = true;
            }
            if (succ) {
                if (!pending && lane == 0) { decided[y] = 1; result[y] = a0 + first_succ; }
                done = true;
            }
        }
        if (!done && !pending && lane == 0) decided[y] = 1;    // every test failed: keep ref (result[y] == ref[y])
    }
    // Everything the host reads after the screening levels in ONE buffer (one D2H instead of four): result, ref, the length of this
    // work list and of the other one. Single-block launches only (the block's warps cover all M responses before the barrier).
    if (summ) {
        __syncthreads();
        for (int y = threadIdx.x; y < M; y += blockDim.x) { summ[y] = result[y]; summ[M + y] = ref[y]; }
        if (threadIdx.x == 0) { summ[2 * M] = work[0]; summ[2 * M + 1] = other_work[0]; }
    }
}

// ---- level 2 ---------------------------------------------------------------------------------------------------------
#ifndef S2_NB_BINS
#define S2_NB_BINS 8192
#endif
constexpr int S2_NB = S2_NB_BINS;      // fine bins: 64 KB of shared-memory counters per CTA at 8192, two CTAs per SM
constexpr int S2_NB_SPLIT = 4096;      // fine bins of the row-split variant (short work lists: every slice merges all its bins)
constexpr size_t S2_SMEM = (size_t)2 * S2_NB * sizeof(uint32_t), S2_SMEM_SPLIT = (size_t)2 * S2_NB_SPLIT * sizeof(uint32_t);
constexpr int S2_THREADS = 512;
constexpr int S2_SPLIT_TESTS = 64;    // work lists up to this long are split over rows
constexpr int S2_MAX_SPLIT = 16;

// Level 2 keeps what the exact level needs of a test it leaves ambiguous: the first rank of each of its 4096 fine bins (XS_CAP
// slots, handed out by an atomic counter; TestInfo::pad holds the slot, ~0u when there was none left).
constexpr int XS_CAP = 128;
constexpr int XS_STRIDE = S2_NB + 1;  // fbase[slot][b] = rank before the first element of fine bin b; [nb] = n (nb = the test's fine bins)

// fine bins per coarse bin, proportional to its population (>= 1): nearly equal-mass fine bins. One warp.
__device__ __forceinline__ void s2_fine_table(const TestInfo* ti, int lane, int nb, uint32_t* off, uint32_t* fc) {
    const uint32_t m0 = ti->pos[2 * lane] + ti->neg[2 * lane], m1 = ti->pos[2 * lane + 1] + ti->neg[2 * lane + 1];
    uint32_t tsum = m0 + m1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tsum += __shfl_xor_sync(0xffffffffu, tsum, o);
    const double share = (tsum > 0) ? (double)(nb - S1_NB) / (double)tsum : 0.0;
    const uint32_t f0 = 1u + (uint32_t)((double)m0 * share), f1 = 1u + (uint32_t)((double)m1 * share);
    uint32_t incl = f0 + f1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    const uint32_t base = incl - f0 - f1;
    off[2 * lane] = base; fc[2 * lane] = f0; off[2 * lane + 1] = base + f0; fc[2 * lane + 1] = f1;
}

// SPLIT = false: one CTA per test over all rows (work lists longer than S2_SPLIT_TESTS); SPLIT = true: the short lists.
// Two instantiations launched back to back, each returning at once when the list is not its kind: the row loop of the
// unsplit kernel is sensitive to code generation (0.47 -> 0.68 ms at C3 with run-time row bounds).
template <bool SPLIT>
__global__ void __launch_bounds__(S2_THREADS, 2) screen2_kernel(const double* __restrict__ T, int64_t ldt, const double* __restrict__ Y, int64_t ldy,
                                                             int64_t n, int M, int A, const double* __restrict__ chk, int nchk, int64_t ldn,
                                                             const double* __restrict__ Q, const double* __restrict__ Eref, double alpha,
                                                             const int* __restrict__ work, TestInfo* __restrict__ info,
                                                             int* __restrict__ status, uint32_t* __restrict__ split_hist,
                                                             unsigned int* __restrict__ split_ticket, unsigned int* __restrict__ xs_count,
                                                             uint32_t* __restrict__ xs_nb, uint32_t* __restrict__ xs_fbase) {
    constexpr int NB = SPLIT ? S2_NB_SPLIT : S2_NB;
    constexpr int BPT = NB / S2_THREADS;
    extern __shared__ __align__(16) uint32_t s2_bins[];
    uint32_t* const pos = s2_bins;
    uint32_t* const neg = s2_bins + NB;
    __shared__ uint32_t off[S1_NB], fc[S1_NB];
    __shared__ long long lred[2][S2_THREADS / 32];
    __shared__ uint32_t wtot[S2_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    __shared__ int s_last2, s_slot;
    const int count = work[0];
    // Short work lists (small sets: 7 tests at C2) would leave one CTA per test streaming all rows alone. Then a test is split
    // over `ns` CTAs by rows; the slices add their non-empty bins into a global histogram of the test and the last slice to
    // arrive (ticket) evaluates the bracket. split_hist / split_ticket are zeroed by the host and hold S2_SPLIT_TESTS tests.
    if ((count <= S2_SPLIT_TESTS) != SPLIT || count <= 0) return;
    const int ns = SPLIT ? max(1, min((int)gridDim.x / count, S2_MAX_SPLIT)) : 1;
    const int nwork = SPLIT ? count * ns : count;
    for (int wj = blockIdx.x; wj < nwork; wj += gridDim.x) {
        const int wi = SPLIT ? wj / ns : wj, slice = SPLIT ? wj - wi * ns : 0;
        const int64_t rows_per = SPLIT ? ((n + ns - 1) / ns + S2_THREADS - 1) / S2_THREADS * S2_THREADS : n;
        const int64_t rbeg = SPLIT ? (int64_t)slice * rows_per : 0, rend = SPLIT ? min(n, rbeg + rows_per) : n;
        const int test = work[1 + wi];
        const int y = test / A, alt = test - y * A;
        TestInfo* ti = info + test;
        __syncthreads();
        for (int i = tid; i < NB; i += S2_THREADS) { pos[i] = 0; neg[i] = 0; }
        if (tid < 32) s2_fine_table(ti, lane, NB, off, fc);
        __syncthreads();
        const double scale = ti->scale;
        const int ncomp = alt + 1;
        const int k = min(ncomp / CHK_G, nchk - 1);
        const double* e0p = (k == 0) ? Y + (int64_t)y * ldy : chk + ((int64_t)y * (nchk - 1) + k - 1) * ldn;
        const double* erp = Eref + (int64_t)y * ldn;
        const int cbeg = k * CHK_G, nfma = ncomp - cbeg;
        double qy[CHK_G];
#pragma unroll
        for (int j = 0; j < CHK_G; j++) qy[j] = (cbeg + j < ncomp) ? Q[(int64_t)(cbeg + j) * M + y] : 0.0;
        const double* tp = T + (int64_t)cbeg * ldt;
        const double* tc[CHK_G - 1];
#pragma unroll
        for (int j = 0; j < CHK_G - 1; j++) tc[j] = tp + (int64_t)min(j, max(nfma - 1, 0)) * ldt;
        auto bin_one = [&](double e, const double er, const double (&tv)[CHK_G - 1]) {
#pragma unroll
            for (int j = 0; j < CHK_G - 1; j++) e = fma(-tv[j], qy[j], e);
            const double d = fabs(er) - fabs(e);
            if (d == 0.0) return;
            // monotone two-level map: coarse bin b (as in level 1), then the position inside it
            const double u = fmin(fabs(d) * scale, (double)S1_NB);       // monotone in |d|
            const int b = min((int)u, S1_NB - 1);
            const double frac = u - (double)b;                           // exact; in [0, 1] (1 only when clamped)
            const uint32_t f = fc[b];
            const uint32_t subi = min((uint32_t)(frac * (double)f), f - 1u);
            atomicAdd((d > 0.0) ? &pos[off[b] + subi] : &neg[off[b] + subi], 1u);
        };
        // Two rows per trip (ten loads in flight per thread): the one-row loop left the warps on the long scoreboard (11.6 warps per
        // issue, ncu r02). Four rows per trip needed 86 registers, one 512-thread CTA per SM instead of two, and was slower; the
        // launch bound keeps two CTAs resident.
        // (An L2 prefetch four trips ahead, which pays in level 1, made this kernel 4 % slower: two CTAs of 16 warps already cover it.)
        constexpr int S2_U = 2;
        int64_t i = rbeg + tid;
        for (; i + (int64_t)(S2_U - 1) * S2_THREADS < rend; i += (int64_t)S2_U * S2_THREADS) {
            double e[S2_U], er[S2_U], tv[S2_U][CHK_G - 1];
#pragma unroll
            for (int u = 0; u < S2_U; u++) {
                const int64_t iu = i + (int64_t)u * S2_THREADS;
                e[u] = e0p[iu]; er[u] = erp[iu];
#pragma unroll
                for (int j = 0; j < CHK_G - 1; j++) tv[u][j] = tc[j][iu];
            }
#pragma unroll
            for (int u = 0; u < S2_U; u++) bin_one(e[u], er[u], tv[u]);
        }
        for (; i < rend; i += S2_THREADS) {
            double tv[CHK_G - 1];
#pragma unroll
            for (int j = 0; j < CHK_G - 1; j++) tv[j] = tc[j][i];
            bin_one(e0p[i], erp[i], tv);
        }
        __syncthreads();
        if (SPLIT && ns > 1) {
            uint32_t* gh = split_hist + (size_t)wi * 2 * NB;
            for (int i = tid; i < NB; i += S2_THREADS) {
                if (pos[i]) atomicAdd(&gh[i], pos[i]);
                if (neg[i]) atomicAdd(&gh[NB + i], neg[i]);
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) s_last2 = (atomicAdd(&split_ticket[wi], 1u) == (unsigned)ns - 1) ? 1 : 0;
            __syncthreads();
            if (!s_last2) continue;                    // uniform over the CTA
            __threadfence();
            for (int i = tid; i < NB; i += S2_THREADS) { pos[i] = __ldcg(&gh[i]); neg[i] = __ldcg(&gh[NB + i]); }
            __syncthreads();
        }
        // exclusive scan of bin populations: thread t owns bins [t*BPT, (t+1)*BPT)
        uint32_t c = 0;
#pragma unroll
        for (int j = 0; j < BPT; j++) c += pos[tid * BPT + j] + neg[tid * BPT + j];
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) wtot[wid] = incl;
        __syncthreads();
        uint32_t wbase = 0;
        for (int ww = 0; ww < wid; ww++) wbase += wtot[ww];
        long long R = (long long)ti->zeros + (long long)(wbase + incl - c);   // zeros hold the lowest ranks
        long long dlo = 0, dhi = 0;
#pragma unroll
        for (int j = 0; j < BPT; j++) {
            const long long pb = pos[tid * BPT + j], nb = neg[tid * BPT + j];
            pos[tid * BPT + j] = (uint32_t)R;              // first rank of the bin, for the exact level (own bins only)
            if (pb + nb) { bin_bounds(R, pb, nb, dlo, dhi); R += pb + nb; }
        }
        dlo = warp_sum_ll(dlo); dhi = warp_sum_ll(dhi);
        if (lane == 0) { lred[0][wid] = dlo; lred[1][wid] = dhi; }
        __syncthreads();
        if (tid == 0) {
            long long lo = 0, hi = 0;
            for (int ww = 0; ww < S2_THREADS / 32; ww++) { lo += lred[0][ww]; hi += lred[1][ww]; }
            const int st = status_from_bounds(lo, hi, (unsigned long long)n, alpha);
            status[test] = st;
            int slot = -1;
            if (st == 2) {          // still ambiguous: keep the fine bins' first ranks for the exact level
                slot = (int)atomicAdd(xs_count, 1u);
                if (slot >= XS_CAP) slot = -1;
                ti->pad = (unsigned int)slot;
                if (slot >= 0) xs_nb[slot] = (uint32_t)NB;
            }
            s_slot = slot;
        }
        __syncthreads();
        if (s_slot >= 0) {
            uint32_t* fb = xs_fbase + (size_t)s_slot * XS_STRIDE;
            for (int i = tid; i < NB; i += S2_THREADS) fb[i] = pos[i];
            if (tid == 0) fb[NB] = (uint32_t)n;
        }
    }
}

// ---- level 3: exact ----------------------------------------------------------------------------------------------------
// keys of work-list entries [w0, w0 + nseg): segment s holds the n keys of test work[1 + w0 + s]
__global__ void work_keys_kernel(const double* __restrict__ T, int64_t ldt, const double* __restrict__ Y, int64_t ldy, int64_t n, int M, int A,
                                 const double* __restrict__ chk, int nchk, int64_t ldn, const double* __restrict__ Q,
                                 const double* __restrict__ Eref, const int* __restrict__ work, int w0, uint64_t* __restrict__ keys,
                                 long long* __restrict__ dsum) {
    const int seg = blockIdx.y;
    const int test = work[1 + w0 + seg];
    const int y = test / A, alt = test - y * A;
    if (blockIdx.x == 0 && threadIdx.x == 0) dsum[seg] = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        keys[(int64_t)seg * n + i] = wilcoxon_key(Eref[(int64_t)y * ldn + i], residual_at(T, ldt, Y, ldy, chk, nchk, ldn, Q, M, y, i, alt + 1));
}

// d = sum_pos (pos+1) * sign  over a sorted segment (exact integer arithmetic; pls.cpp:202)
__global__ void __launch_bounds__(256) ranksum_kernel(const uint64_t* __restrict__ keys, int64_t n, long long* __restrict__ dsum) {
    const int seg = blockIdx.y;
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
        atomicAdd((unsigned long long*)&dsum[seg], (unsigned long long)t);   // integer: order independent
    }
}

// ---- level 3 from the fine bins of level 2 -------------------------------------------------------------------------------
// A test that level 2 leaves ambiguous has every element's rank pinned to its fine bin (~n / 4096 elements) and the first rank of
// every fine bin on record (xs_fbase). So instead of sorting the test's n keys (eight radix passes), its elements are scattered
// into their bins in one pass (exact_scatter_kernel: positives fill a bin from the front, negatives from the back, in whatever
// order the atomics yield) and d follows from counting inside the bins (exact_rank_kernel). With p positives, q negatives,
// cnt = p + q elements in a bin whose first rank is R + 1, and r_i the rank of element i inside the bin (0 .. cnt - 1):
//     sum_i s_i (R + 1 + r_i) = (R + 1)(p - q) + 2 sum_{i positive} r_i - cnt (cnt - 1) / 2,
//     sum_{i positive} r_i    = p (p - 1) / 2 + C,     C = #{(i positive, j negative) : key_j < key_i}
// (equal keys carry the same sign, so ties never cross signs and who takes which rank among them does not change d; at equal |d| a
// negative sorts below a positive, as in the key order of the radix path). Only C needs the keys: p q comparisons per bin instead
// of a sort, none at all for the bins that hold one sign only. d is summed in integers.
// The scatter recomputes the differences and bins exactly as screen2_kernel does (same expressions, same checkpoint and padded
// FMAs); should a bin ever receive more elements than level 2 counted, or hold more than XS_BIG, a flag sends the call to the
// radix path below, which does not depend on any of this.
constexpr int XS_THREADS = 256;
constexpr int XS_BIG = 32768;         // elements per fine bin the counting accepts (16-bit cursors per sign; p q comparisons)
constexpr int XR_THREADS = 256;
constexpr int XR_CTAS = 32;           // CTAs per test in exact_rank_kernel

// xs_cursor[slot][bin]: positives so far in the low half, negatives so far in the high half
__global__ void __launch_bounds__(XS_THREADS) exact_scatter_kernel(const double* __restrict__ T, int64_t ldt, const double* __restrict__ Y, int64_t ldy,
                                                                  int64_t n, int M, int A, const double* __restrict__ chk, int nchk, int64_t ldn,
                                                                  const double* __restrict__ Q, const double* __restrict__ Eref,
                                                                  const int* __restrict__ work, int w0, const TestInfo* __restrict__ info,
                                                                  const uint32_t* __restrict__ xs_nb, const uint32_t* __restrict__ xs_fbase,
                                                                  uint32_t* __restrict__ xs_cursor, unsigned int* __restrict__ xs_flag,
                                                                  uint64_t* __restrict__ keys) {
    __shared__ uint32_t off[S1_NB], fc[S1_NB];
    const int seg = blockIdx.y, tid = threadIdx.x;
    const int test = work[1 + w0 + seg];
    const int y = test / A, alt = test - y * A;
    const TestInfo* ti = info + test;
    const unsigned int slot = ti->pad;
    if (slot >= (unsigned)XS_CAP) { if (tid == 0 && blockIdx.x == 0) atomicOr(xs_flag, 1u); return; }   // level 2 had no slot left
    if (tid < 32) s2_fine_table(ti, tid, (int)xs_nb[slot], off, fc);
    __syncthreads();
    const double scale = ti->scale;
    const int ncomp = alt + 1;
    const int k = min(ncomp / CHK_G, nchk - 1);
    const double* e0p = (k == 0) ? Y + (int64_t)y * ldy : chk + ((int64_t)y * (nchk - 1) + k - 1) * ldn;
    const double* erp = Eref + (int64_t)y * ldn;
    const int cbeg = k * CHK_G, nfma = ncomp - cbeg;
    double qy[CHK_G - 1];
    const double* tc[CHK_G - 1];
#pragma unroll
    for (int j = 0; j < CHK_G - 1; j++) {
        qy[j] = (cbeg + j < ncomp) ? Q[(int64_t)(cbeg + j) * M + y] : 0.0;
        tc[j] = T + (int64_t)cbeg * ldt + (int64_t)min(j, max(nfma - 1, 0)) * ldt;
    }
    const uint32_t* fb = xs_fbase + (size_t)slot * XS_STRIDE;
    uint32_t* cur = xs_cursor + (size_t)slot * S2_NB;
    uint64_t* kout = keys + (int64_t)seg * n;
    for (int64_t i = (int64_t)blockIdx.x * XS_THREADS + tid; i < n; i += (int64_t)gridDim.x * XS_THREADS) {
        double e = e0p[i];
        const double er = erp[i];
#pragma unroll
        for (int j = 0; j < CHK_G - 1; j++) e = fma(-tc[j][i], qy[j], e);
        const double d = fabs(er) - fabs(e);
        if (d == 0.0) continue;                                            // zeros: lowest ranks, sign 0 (pls.cpp:196)
        const double u = fmin(fabs(d) * scale, (double)S1_NB);             // the map of screen2_kernel, expression by expression
        const int b = min((int)u, S1_NB - 1);
        const double frac = u - (double)b;
        const uint32_t f = fc[b];
        const uint32_t subi = min((uint32_t)(frac * (double)f), f - 1u);
        const uint32_t fbin = off[b] + subi;
        const bool posv = d > 0.0;
        const uint32_t old = atomicAdd(&cur[fbin], posv ? 1u : 0x10000u);
        const uint32_t pc = old & 0xffffu, qc = old >> 16;
        const uint32_t first = fb[fbin], cnt = fb[fbin + 1] - first;
        if (pc + qc < cnt) kout[posv ? first + pc : first + cnt - 1u - qc] = ((uint64_t)__double_as_longlong(fabs(d)) << 1) | (uint64_t)posv;
        else atomicOr(xs_flag, 2u);
    }
}

// grid (XR_CTAS, tests). Work unit = (fine bin with both signs, 32 of its positives) against all of the bin's negatives; the warps of
// all CTAs of a test take the units round robin, so the few crowded bins (the clamped last coarse bin holds the whole tail of the
// distribution) are spread over all of them. CTA 0 of a test adds the terms that need no keys.
__global__ void __launch_bounds__(XR_THREADS) exact_rank_kernel(const uint64_t* __restrict__ keys, int64_t n, const int* __restrict__ work, int w0,
                                                               const TestInfo* __restrict__ info, const uint32_t* __restrict__ xs_nb,
                                                               const uint32_t* __restrict__ xs_fbase, const uint32_t* __restrict__ xs_cursor,
                                                               unsigned int* __restrict__ xs_flag, long long* __restrict__ dsum) {
    __shared__ uint32_t us[S2_NB + 1];        // first work unit of each bin; [S2_NB] = number of units
    __shared__ uint32_t wtot[XR_THREADS / 32];
    __shared__ long long red[XR_THREADS / 32];
    const int seg = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned int slot = info[work[1 + w0 + seg]].pad;
    if (slot >= (unsigned)XS_CAP) return;
    const uint32_t* fb = xs_fbase + (size_t)slot * XS_STRIDE;
    const uint32_t* pq = xs_cursor + (size_t)slot * S2_NB;       // positives | negatives << 16 of each bin, as the scatter left them
    const int nb = (int)xs_nb[slot], bpt = nb / XR_THREADS;      // fine bins of this test (a multiple of XR_THREADS), bins per thread
    long long acc = 0;
    uint32_t c = 0;
    bool bad = false;
#pragma unroll 4
    for (int j = 0; j < bpt; j++) {
        const int b = tid * bpt + j;
        const uint32_t first = fb[b], cnt = fb[b + 1] - first;
        const uint32_t pb = pq[b] & 0xffffu, qb = pq[b] >> 16;
        bad |= (pb + qb != cnt) || cnt > (uint32_t)XS_BIG;
        if (pb && qb) c += (pb + 31u) >> 5;
        if (blockIdx.x == 0)
            acc += (long long)(first + 1u) * ((long long)pb - (long long)qb) + (long long)pb * (long long)(pb - 1u) - (long long)cnt * (long long)(cnt - 1u) / 2;
    }
    if (bad && blockIdx.x == 0) atomicOr(xs_flag, 4u);
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) wtot[wid] = incl;
    __syncthreads();
    uint32_t run = incl - c;
    for (int ww = 0; ww < wid; ww++) run += wtot[ww];
#pragma unroll 4
    for (int j = 0; j < bpt; j++) {
        const int b = tid * bpt + j;
        us[b] = run;
        const uint32_t pb = pq[b] & 0xffffu, qb = pq[b] >> 16;
        if (pb && qb) run += (pb + 31u) >> 5;
    }
    if (tid == XR_THREADS - 1) us[nb] = run;
    __syncthreads();
    const uint32_t U = us[nb];
    const uint64_t* kk = keys + (int64_t)seg * n;
    for (uint32_t u = blockIdx.x * (XR_THREADS / 32) + wid; u < U; u += gridDim.x * (XR_THREADS / 32)) {
        int lo = 0, hi = nb;                    // smallest index with us[index] > u; the unit's bin is the one before it
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (us[mid] > u) hi = mid; else lo = mid + 1; }
        const int b = lo - 1;
        const uint32_t first = fb[b], cnt = fb[b + 1] - first;
        const uint32_t pb = pq[b] & 0xffffu;
        const uint32_t i = (u - us[b]) * 32u + lane;
        const uint64_t* kb = kk + first;
        const uint64_t ki = (i < pb) ? kb[i] : 0ull;     // a positive of the bin (0 is below every key of a nonzero difference)
        uint32_t below = 0;
#pragma unroll 4
        for (uint32_t j = pb; j < cnt; j++) below += (kb[j] < ki) ? 1u : 0u;     // the bin's negatives
        acc += 2ll * (long long)below;
    }
    acc = warp_sum_ll(acc);
    if (lane == 0) red[wid] = acc;
    __syncthreads();
    if (tid == 0) {
        long long t = 0;
        for (int w = 0; w < XR_THREADS / 32; w++) t += red[w];
        if (t) atomicAdd((unsigned long long*)&dsum[seg], (unsigned long long)t);       // integer: order independent
    }
}

// skip != null: nothing is written when *skip != 0 (the fine-bin level gave up; the radix path will redo these tests)
__global__ void exact_status_kernel(const long long* __restrict__ dsum, const int* __restrict__ work, int w0, int nseg, unsigned long long n,
                                    double alpha, int* __restrict__ status, const unsigned int* __restrict__ skip) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    if (skip && *skip) return;
    status[work[1 + w0 + s]] = (wilcoxon_p_from_d(dsum[s], n) > alpha) ? 1 : 0;
}

__global__ void single_keys_kernel(const double* __restrict__ e1, const double* __restrict__ e2, int64_t n, uint64_t* __restrict__ keys) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        keys[i] = wilcoxon_key(e1[i], e2[i]);
}
__global__ void single_p_kernel(const long long* __restrict__ dsum, unsigned long long n, double* __restrict__ p) {
    *p = wilcoxon_p_from_d(*dsum, n);
}

constexpr size_t S2_SPLIT_BYTES = ((size_t)S2_SPLIT_TESTS * 2 * S2_NB_SPLIT + S2_SPLIT_TESTS) * sizeof(uint32_t);
// xs: [0] slots handed out, [1] flags of the fine-bin exact level, [16 + slot] fine bins of the slot's test, then the cursors and
// the first ranks
constexpr size_t XS_HEAD = 16 + XS_CAP;
constexpr size_t XS_BYTES = (XS_HEAD + (size_t)XS_CAP * S2_NB + (size_t)XS_CAP * XS_STRIDE) * sizeof(uint32_t);

struct HoldPlan { int nchk; int64_t ldn; int nblk; int64_t rows_per_blk; int ycta; int exact_cap; int ngroup, nsplit; int64_t rows_per_split; };

HoldPlan hold_plan(const abcb200_ctx* ctx, int64_t n_te, int M, int A) {
    HoldPlan p;
    p.nchk = (A + CHK_G - 1) / CHK_G;                  // checkpoints 0 (= Y) .. nchk-1
    p.ldn = (n_te + 31) / 32 * 32;
    p.ycta = (M + PC_MY - 1) / PC_MY;
    p.rows_per_blk = PC_ROWS;
    p.nblk = (int)((n_te + PC_ROWS - 1) / PC_ROWS);
    int64_t cap = (int64_t)(1.0e9 / (16.0 * (double)n_te));           // keys + alt buffer <= ~1 GB
    p.exact_cap = (int)max((int64_t)1, min(cap, (int64_t)256));
    // level 1: groups of S1_TESTS tests per response; small grids are split over rows to fill the machine
    p.ngroup = max(1, (A - 1 + S1_TESTS - 1) / S1_TESTS);
    const int64_t trip = (int64_t)S1_ROWS * S1_THREADS;
    // aim at ~8 waves of the resident capacity (the groups above ref[y] exit at once, so the live grid is smaller), but keep
    // >= 4096 rows per split: every CTA pays for clearing its counters, the scale sample and the merge of 516 totals
    const int64_t cap_ctas = 8 * (int64_t)S1_CTAS_PER_SM * ctx->sm_count, live = (int64_t)p.ngroup * M;
    int64_t ns = (cap_ctas + live - 1) / live;
    ns = max((int64_t)1, min(ns, (n_te + 4095) / 4096));
    int64_t rps = (n_te + ns - 1) / ns;
    rps = (rps + trip - 1) / trip * trip;
    p.rows_per_split = rps;
    p.nsplit = (int)((n_te + rps - 1) / rps);
    return p;
}

}  // namespace

size_t holdout_ws_bytes(const abcb200_ctx* ctx, int64_t n_te, int K, int M, int A, bool own_scores) {
    if (n_te <= 0) return 4096;
    const HoldPlan p = hold_plan(ctx, n_te, M, A);
    size_t b = 0;
    if (own_scores) b += align_up((size_t)p.ldn * A * 8, 256);                          // T
    b += align_up((size_t)p.nblk * M * A * 8, 256);                                    // PRESS partials
    b += align_up((size_t)M * A * 8, 256);                                             // PRESS
    b += align_up((size_t)max(1, p.nchk - 1) * M * p.ldn * 8, 256);                     // checkpoints
    b += align_up((size_t)M * p.ldn * 8, 256);                                         // Eref
    b += 3 * align_up((size_t)M * 4, 256);                                             // ref, decided, result
    b += align_up((size_t)M * A * 4, 256);                                             // status
    b += 2 * align_up(((size_t)M * A + 1) * 4, 256);                                   // work lists
    b += align_up((size_t)M * A * sizeof(TestInfo), 256);
    b += align_up((size_t)M * p.ngroup * (S1_GH + 1) * 4, 256);                        // level-1 merged totals + tickets
    b += align_up(S2_SPLIT_BYTES, 256);                                                // level-2 split histograms + tickets
    b += align_up(XS_BYTES, 256);                                                      // level-2 fine-bin ranks kept for the exact level
    b += align_up((2 * (size_t)M + 2) * 4, 256);                                       // host summary
    b += 2 * align_up((size_t)p.exact_cap * n_te * 8, 256);                            // keys, keys_alt
    b += radix_hist_bytes(n_te, p.exact_cap);
    b += align_up((size_t)p.exact_cap * 8, 256);
    return b + 8192;
}

// Buffers of one hold-out validation. T_ext: hold-out scores the caller produces itself (n_te x A, ld ldt_ext; the pipelined
// ranking fills them block by block), or null: an own buffer, filled by holdout_scores().
int holdout_begin(abcb200_ctx* ctx, const double* Yte, int64_t ldy, int64_t n_te, const PlsFactors& f, const double* T_ext, int64_t ldt_ext,
                  double* press_dev, HoldoutJob* job) {
    const int M = f.M, A = f.A;
    if (n_te > 0x7fffffffll) ABC_FAIL(ctx, ABCB200_EINVAL, "holdout: %lld hold-out rows exceed the 32-bit counters", (long long)n_te);
    if (M > 128) ABC_FAIL(ctx, ABCB200_EINVAL, "holdout: M=%d responses exceed 128 (one-block summary of the selection)", M);
    const HoldPlan p = hold_plan(ctx, n_te, M, A);
    HoldoutJob& j = *job;
    j.n_te = n_te; j.K = f.K; j.M = M; j.A = A; j.Yte = Yte; j.ldy = ldy; j.Q = f.Q;
    j.nchk = p.nchk; j.ldn = p.ldn; j.nblk = p.nblk; j.ycta = p.ycta; j.exact_cap = p.exact_cap; j.ngroup = p.ngroup; j.nsplit = p.nsplit;
    j.rows_per_split = p.rows_per_split;
    j.ldt = T_ext ? ldt_ext : p.ldn;
    j.T = T_ext ? const_cast<double*>(T_ext) : ws_new<double>(ctx, (size_t)p.ldn * A);
    j.partial = ws_new<double>(ctx, (size_t)p.nblk * M * A);
    j.press = press_dev ? press_dev : ws_new<double>(ctx, (size_t)M * A);
    j.chk = ws_new<double>(ctx, (size_t)max(1, p.nchk - 1) * M * p.ldn);
    j.Eref = ws_new<double>(ctx, (size_t)M * p.ldn);
    j.ref = ws_new<int>(ctx, M);
    j.decided = ws_new<int>(ctx, M);
    j.result = ws_new<int>(ctx, M);
    j.status = ws_new<int>(ctx, (size_t)M * A);
    j.work1 = ws_new<int>(ctx, (size_t)M * A + 1);
    j.work2 = ws_new<int>(ctx, (size_t)M * A + 1);
    j.info = ws_alloc(ctx, (size_t)M * A * sizeof(TestInfo));
    j.ghist = ws_new<unsigned int>(ctx, (size_t)M * p.ngroup * (S1_GH + 1));
    j.ticket = j.ghist ? j.ghist + (size_t)M * p.ngroup * S1_GH : nullptr;
    j.s2hist = (uint32_t*)ws_alloc(ctx, S2_SPLIT_BYTES);
    j.xs = (uint32_t*)ws_alloc(ctx, XS_BYTES);
    j.summ = ws_new<int>(ctx, 2 * (size_t)M + 2);
    if (!j.summ || !j.s2hist || !j.xs || !j.ghist || !j.T || !j.partial || !j.press || !j.chk || !j.Eref || !j.ref || !j.decided || !j.result || !j.status || !j.work1 ||
        !j.work2 || !j.info)
        ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in holdout_select");
    CUDA_TRY(ctx, cudaFuncSetAttribute(press_chk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PC_SMEM));
    return ABCB200_OK;
}

// own hold-out scores, all A components at once: T = Zte R (kernel timer 5)
int holdout_scores(abcb200_ctx* ctx, const HoldoutJob* job, const double* Zte, int64_t ldx, const double* R) {
    kernel_begin(ctx, 5);
    ABC_TRY(launch_xb(ctx, Zte, ldx, job->n_te, job->K, R, job->K, job->A, job->T, job->ldt));
    kernel_end(ctx, 5);
    return ABCB200_OK;
}

// PRESS partial sums and checkpoints of components [c_begin, c_end) (c_begin a multiple of CHK_G; c_end a multiple of it or A);
// the scores of those components must be in job->T. Kernel timer 4.
int holdout_press_block(abcb200_ctx* ctx, const HoldoutJob* job, int c_begin, int c_end) {
    const HoldoutJob& j = *job;
    const int k_begin = c_begin / CHK_G, k_end = (c_end + CHK_G - 1) / CHK_G;
    if (k_end <= k_begin) return ABCB200_OK;
    kernel_begin(ctx, 4);
    LAUNCH(ctx, press_chk_kernel, (unsigned)((int64_t)j.nblk * j.ycta), PC_THREADS, PC_SMEM, j.T, j.ldt, j.Yte, j.ldy, j.n_te, j.M, j.A, j.Q, j.ycta, j.nchk, j.ldn, j.chk,
           j.partial, k_begin, k_end);
    kernel_end(ctx, 4);
    return ABCB200_OK;
}

// PRESS, its first arg-min per response and the selection state (the end of stage 2)
int holdout_press_finalize(abcb200_ctx* ctx, const HoldoutJob* job) {
    const HoldoutJob& j = *job;
    LAUNCH(ctx, press_reduce_kernel, dim3(j.M, (j.A + PF_CW - 1) / PF_CW), PF_T, 0, j.partial, j.nblk, j.M, j.A, j.press);
    LAUNCH(ctx, press_argmin_kernel, (j.M + 3) / 4, 128, 0, j.press, j.M, j.A, j.ref, j.decided, j.result);
    return ABCB200_OK;
}

// Stage 3: component selection (pls.cpp:265-289) from the finished PRESS / checkpoints / scores. ncomp_host: M counts.
int holdout_select_finish(abcb200_ctx* ctx, const HoldoutJob* job, double alpha, int32_t* ncomp_host) {
    const HoldoutJob& p = *job;
    const int M = p.M, A = p.A;
    const int64_t n_te = p.n_te, ldt = p.ldt, ldy = p.ldy;
    const double *T = p.T, *Yte = p.Yte, *Q = p.Q;
    double *chk = p.chk, *Eref = p.Eref;
    int *ref = p.ref, *decided = p.decided, *result = p.result, *status = p.status, *work1 = p.work1, *work2 = p.work2, *summ = p.summ;
    TestInfo* info = (TestInfo*)p.info;
    unsigned int *ghist = p.ghist, *ticket = p.ticket;
    uint32_t* s2hist = p.s2hist;
    uint32_t *xs_cursor = p.xs + XS_HEAD, *xs_fbase = xs_cursor + (size_t)XS_CAP * S2_NB;
    unsigned int *xs_count = p.xs, *xs_flag = p.xs + 1;
    uint32_t* xs_nb = p.xs + 16;
    stage_begin(ctx, 3);
    const int egrid = (int)max((int64_t)1, min((n_te + 255) / 256, (int64_t)(4 * ctx->sm_count)));
    LAUNCH(ctx, eref_kernel, dim3(egrid, M), 256, 0, T, ldt, Yte, ldy, n_te, M, chk, p.nchk, p.ldn, Q, ref, Eref);
    CUDA_TRY(ctx, cudaMemsetAsync(work1, 0, sizeof(int), ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(work2, 0, sizeof(int), ctx->stream));
    if (A > 1) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(screen1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S1_SMEM));
        CUDA_TRY(ctx, cudaFuncSetAttribute(screen1_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        if (p.nsplit > 1) CUDA_TRY(ctx, cudaMemsetAsync(ghist, 0, (size_t)M * p.ngroup * (S1_GH + 1) * 4, ctx->stream));
        static const int s1_pf = getenv("ABCB200_S1_PF") ? atoi(getenv("ABCB200_S1_PF")) : S1_PF;      // tuning knob: L2 prefetch distance in trips
        kernel_begin(ctx, 2);
        LAUNCH(ctx, screen1_kernel, dim3(M, p.ngroup, p.nsplit), S1_THREADS, S1_SMEM, T, ldt, Yte, ldy, n_te, M, A, chk, p.nchk, p.ldn,
               Q, Eref, ref, alpha, p.rows_per_split, ghist, ticket, status, info, s1_pf);
        kernel_end(ctx, 2);
    }
    LAUNCH(ctx, decide_kernel, (M + 3) / 4, 128, 0, status, ref, M, A, decided, result, work1, (int*)nullptr, (const int*)nullptr);
    kernel_begin(ctx, 3);
    CUDA_TRY(ctx, cudaMemsetAsync(s2hist, 0, S2_SPLIT_BYTES, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(p.xs, 0, XS_HEAD * sizeof(uint32_t), ctx->stream));
    CUDA_TRY(ctx, cudaFuncSetAttribute(screen2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S2_SMEM));
    CUDA_TRY(ctx, cudaFuncSetAttribute(screen2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S2_SMEM_SPLIT));
    LAUNCH(ctx, screen2_kernel<false>, 2 * ctx->sm_count, S2_THREADS, S2_SMEM, T, ldt, Yte, ldy, n_te, M, A, chk, p.nchk, p.ldn, Q, Eref, alpha, work1, info, status,
           s2hist, (unsigned int*)(s2hist + (size_t)S2_SPLIT_TESTS * 2 * S2_NB_SPLIT), xs_count, xs_nb, xs_fbase);
    LAUNCH(ctx, screen2_kernel<true>, 2 * ctx->sm_count, S2_THREADS, S2_SMEM_SPLIT, T, ldt, Yte, ldy, n_te, M, A, chk, p.nchk, p.ldn, Q, Eref, alpha, work1, info, status,
           s2hist, (unsigned int*)(s2hist + (size_t)S2_SPLIT_TESTS * 2 * S2_NB_SPLIT), xs_count, xs_nb, xs_fbase);
    kernel_end(ctx, 3);
    LAUNCH(ctx, decide_kernel, 1, 1024, 0, status, ref, M, A, decided, result, work2, summ, (const int*)work1);   // one block (the summary needs every response's result)
    ABC_TRY(hpin_reserve(ctx, sizeof(int) * (2 * (size_t)M + 4) + 64));
    int* h_result = (int*)ctx->hpin;
    int* h_ref = h_result + M;
    int* h_count = h_ref + M;
    CUDA_TRY(ctx, cudaMemcpyAsync(h_result, summ, sizeof(int) * (2 * (size_t)M + 2), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    const int n_exact = *h_count;
    ctx->stat_level2 = (uint64_t)h_count[1];
    ctx->stat_tests = 0;
    for (int y = 0; y < M; y++) ctx->stat_tests += (uint64_t)h_ref[y];
    if (n_exact > 0) {   // intervals still straddling the threshold after level 2: rank exactly those tests
        uint64_t* keys = ws_new<uint64_t>(ctx, (size_t)p.exact_cap * n_te);
        uint64_t* keys_alt = ws_new<uint64_t>(ctx, (size_t)p.exact_cap * n_te);
        uint32_t* hist = (uint32_t*)ws_alloc(ctx, radix_hist_bytes(n_te, p.exact_cap));
        long long* dsum = ws_new<long long>(ctx, p.exact_cap);
        if (!keys || !keys_alt || !hist || !dsum) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in holdout_select (exact tests)");
        const int kgrid = (int)max((int64_t)1, min((n_te + 255) / 256, (int64_t)(2 * ctx->sm_count)));
        const int rgrid = (int)max((int64_t)1, min((n_te + 2047) / 2048, (int64_t)64));
        bool radix = getenv("ABCB200_EXACT_RADIX") != nullptr;      // force the radix path (tests, A/B timing); read per call on purpose
        if (!radix) {    // from the fine bins of level 2 (one scatter pass + ranking inside the bins)
            CUDA_TRY(ctx, cudaMemsetAsync(xs_cursor, 0, (size_t)XS_CAP * S2_NB * sizeof(uint32_t), ctx->stream));
            unsigned int* h_flag = (unsigned int*)(h_count + 2);
            for (int w0 = 0; w0 < n_exact; w0 += p.exact_cap) {
                const int nseg = min(p.exact_cap, n_exact - w0);
                const int sgrid = (int)max((int64_t)1, min((n_te + 4 * XS_THREADS - 1) / (4 * XS_THREADS), (int64_t)max(1, 16 * ctx->sm_count / nseg)));
                CUDA_TRY(ctx, cudaMemsetAsync(dsum, 0, sizeof(long long) * nseg, ctx->stream));
                LAUNCH(ctx, exact_scatter_kernel, dim3(sgrid, nseg), XS_THREADS, 0, T, ldt, Yte, ldy, n_te, M, A, chk, p.nchk, p.ldn, Q, Eref, work2, w0,
                       (const TestInfo*)info, (const uint32_t*)xs_nb, (const uint32_t*)xs_fbase, xs_cursor, xs_flag, keys);
                LAUNCH(ctx, exact_rank_kernel, dim3(XR_CTAS, nseg), XR_THREADS, 0, (const uint64_t*)keys, n_te, work2, w0, (const TestInfo*)info,
                       (const uint32_t*)xs_nb, (const uint32_t*)xs_fbase, (const uint32_t*)xs_cursor, xs_flag, dsum);
                LAUNCH(ctx, exact_status_kernel, (nseg + 127) / 128, 128, 0, dsum, work2, w0, nseg, (unsigned long long)n_te, alpha, status, (const unsigned int*)xs_flag);
            }
            CUDA_TRY(ctx, cudaMemsetAsync(work1, 0, sizeof(int), ctx->stream));
            LAUNCH(ctx, decide_kernel, (M + 3) / 4, 128, 0, status, ref, M, A, decided, result, work1, (int*)nullptr, (const int*)nullptr);
            CUDA_TRY(ctx, cudaMemcpyAsync(h_result, result, sizeof(int) * M, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(ctx, cudaMemcpyAsync(h_flag, xs_flag, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            radix = *h_flag != 0;            // no slot left / a crowded bin: nothing was decided from these tests, sort them instead
        }
        if (radix) {
            ctx->exact_radix_calls++;
            for (int w0 = 0; w0 < n_exact; w0 += p.exact_cap) {
                const int nseg = min(p.exact_cap, n_exact - w0);
                LAUNCH(ctx, work_keys_kernel, dim3(kgrid, nseg), 256, 0, T, ldt, Yte, ldy, n_te, M, A, chk, p.nchk, p.ldn, Q, Eref, work2, w0, keys, dsum);
                ABC_TRY(radix_sort_segments(ctx, keys, keys_alt, nullptr, nullptr, n_te, nseg, hist, nullptr));
                LAUNCH(ctx, ranksum_kernel, dim3(rgrid, nseg), 256, 0, keys, n_te, dsum);
                LAUNCH(ctx, exact_status_kernel, (nseg + 127) / 128, 128, 0, dsum, work2, w0, nseg, (unsigned long long)n_te, alpha, status, (const unsigned int*)nullptr);
            }
            CUDA_TRY(ctx, cudaMemsetAsync(work1, 0, sizeof(int), ctx->stream));
            LAUNCH(ctx, decide_kernel, (M + 3) / 4, 128, 0, status, ref, M, A, decided, result, work1, (int*)nullptr, (const int*)nullptr);
            CUDA_TRY(ctx, cudaMemcpyAsync(h_result, result, sizeof(int) * M, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        }
        stage_end(ctx, 3);
    } else {
        stage_end(ctx, 3);
    }
    ctx->exact_tests += (uint64_t)n_exact;
    for (int y = 0; y < M; y++) ncomp_host[y] = h_result[y] + 1;   // index -> component count (pls.cpp:288)
    return ABCB200_OK;
}

int holdout_select_dev(abcb200_ctx* ctx, const double* Zte, int64_t ldx, const double* Yte, int64_t ldy, int64_t n_te,
                       const PlsFactors& f, double alpha, double* press_dev, int32_t* ncomp_host) {
    const int M = f.M, A = f.A;
    if (n_te <= 0) {   // empty hold-out: PRESS all zero -> argmin 0 -> one component for every response
        if (press_dev) CUDA_TRY(ctx, cudaMemsetAsync(press_dev, 0, sizeof(double) * M * A, ctx->stream));
        if (ncomp_host) for (int y = 0; y < M; y++) ncomp_host[y] = 1;
        return ABCB200_OK;
    }
    stage_begin(ctx, 2);
    HoldoutJob job;
    ABC_TRY(holdout_begin(ctx, Yte, ldy, n_te, f, nullptr, 0, press_dev, &job));
    ABC_TRY(holdout_scores(ctx, &job, Zte, ldx, f.R));            // hold-out scores, all A components
    ABC_TRY(holdout_press_block(ctx, &job, 0, A));
    ABC_TRY(holdout_press_finalize(ctx, &job));
    stage_end(ctx, 2);
    if (!ncomp_host) return ABCB200_OK;
    return holdout_select_finish(ctx, &job, alpha, ncomp_host);
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
    LAUNCH(ctx, ranksum_kernel, dim3((int)max((int64_t)1, min((n + 2047) / 2048, (int64_t)64)), 1), 256, 0, keys, n, dsum);
    LAUNCH(ctx, single_p_kernel, 1, 1, 0, dsum, (unsigned long long)n, p);
    ABC_TRY(hpin_reserve(ctx, 64));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hpin, p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *p_host = *(double*)ctx->hpin;
    return ABCB200_OK;
}
