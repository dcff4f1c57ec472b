// api.cu — the C ABI (include/abcsmc_b200.h): argument checks, workspace planning, host<->device staging and
// the per-stage orchestration of the kernels. No numerics live here.
#include <stdlib.h>

#include <algorithm>
#include <cmath>
#include <new>
#include <numeric>
#include <vector>

#include "kernels.cuh"

struct abcb200_pls {
    abcb200_ctx* ctx;
    PlsFactors f;
    double* storage;   // one cudaMalloc holding W,P,R,Q,(T)
};

namespace {

inline int64_t pad32(int64_t n) { return (n + 31) / 32 * 32; }

int check_ctx(abcb200_ctx* ctx) {
    if (!ctx) return ABCB200_EINVAL;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return ABCB200_ENODEV;
    ctx->err[0] = 0;
    return ABCB200_OK;
}

// column-major host -> device copy with leading dimensions
int h2d_matrix(abcb200_ctx* ctx, double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols) {
    if (rows <= 0 || cols <= 0) return ABCB200_OK;
    CUDA_TRY(ctx, cudaMemcpy2DAsync(dst, (size_t)ldd * 8, src, (size_t)lds * 8, (size_t)rows * 8, (size_t)cols, cudaMemcpyHostToDevice, ctx->stream));
    return ABCB200_OK;
}
int d2h(abcb200_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return ABCB200_OK;
    CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return ABCB200_OK;
}

// The pipelined S2 + S3 (rank_fit_holdout_pipelined) applies when there is a hold-out set and the component loop is one of the two
// Gram-based ones: the on-chip loop (pls_defl.cu; lanes = ordinary high / low priority streams) or the wide loop (pls_wide.cu, three
// small launches per component; lanes = an SM partition, because its launches would otherwise queue behind the consumers' CTAs).
// ABCB200_NO_PIPELINE=1 keeps the stage-after-stage order (A/B measurements, debugging); ABCB200_SM_PARTITION=0 refuses partitions
// (the wide shapes then run stage after stage), =2 gives the on-chip loop an 8-SM partition as well.
constexpr int PIPE_WIDE_SMS = 24;
struct PipePlan { int kind; int block; cudaStream_t small, rest; };      // kind 0: not pipelined, 1: on-chip loop, 2: wide loop
PipePlan rank_pipe_plan(const abcb200_ctx* cctx, int K, int P, int method, int64_t n_te) {
    static const bool off = getenv("ABCB200_NO_PIPELINE") != nullptr || getenv("ABCB200_PLS_LITERAL") != nullptr || getenv("ABCB200_PLS_PROF") != nullptr;
    static const int part_env = getenv("ABCB200_SM_PARTITION") ? atoi(getenv("ABCB200_SM_PARTITION")) : 1;
    static const int wide_sms = getenv("ABCB200_WIDE_SMS") ? atoi(getenv("ABCB200_WIDE_SMS")) : PIPE_WIDE_SMS;      // tuning knob (a multiple of 8)
    abcb200_ctx* ctx = const_cast<abcb200_ctx*>(cctx);
    PipePlan pl{0, 0, nullptr, nullptr};
    if (off || method == ABCB200_KERNEL_TYPE1_STREAM || n_te <= 0) return pl;
    if (pls_defl_fits(ctx, K, P)) {
        if (!(part_env >= 2 && ctx_lanes(ctx, 8, &pl.small, &pl.rest))) ctx_lanes(ctx, 0, &pl.small, &pl.rest);
        pl.kind = 1; pl.block = 32;
    } else if (pls_wide_fits(ctx, K, P) && ctx_lanes(ctx, wide_sms, &pl.small, &pl.rest)) {
        pl.kind = 2; pl.block = std::max(64, ((K + 15) / 16 + 31) / 32 * 32);
    }
    return pl;
}
bool rank_pipelined(const abcb200_ctx* ctx, int K, int P, int method, int64_t n_te) { return rank_pipe_plan(ctx, K, P, method, n_te).kind != 0; }

size_t rank_core_ws_bytes(const abcb200_ctx* ctx, int64_t N, int K, int P, double f, int method, bool simple) {
    const int64_t ldz = pad32(N);
    size_t b = 0;
    b += moments_ws_bytes(N, K + P);
    b += align_up((size_t)ldz * K * 8, 256);
    b += 8 * align_up((size_t)(K + P) * 8, 256);
    b += align_up((size_t)N * 8, 256);       // distances
    b += order_ws_bytes(N);
    if (!simple) {
        const int64_t n_tr = (int64_t)std::llround((double)N * f);
        const int64_t n_te = N - n_tr;
        b += align_up((size_t)ldz * P * 8, 256);
        b += 3 * align_up((size_t)K * K * 8, 256) + align_up((size_t)P * K * 8, 256);
        b += pls_fit_ws_bytes(ctx, n_tr, K, P, method);
        b += holdout_ws_bytes(ctx, n_te, K, P, K);
        b += align_up((size_t)K * 8, 256);
        if (rank_pipelined(ctx, K, P, method, n_te))      // scores of ALL rows for all A components + the chunked loop's state
            b += align_up((size_t)(ldz + 64) * K * 8, 256) + align_up(pls_defl_state_doubles(K, P) * 8, 256) + align_up((size_t)K * K * 8, 256) + pls_wide_ws_bytes(K, P, K) + 4096;
    }
    return b + 16384;
}

// dist[i] = || T[i, :ncols] - ref ||_2 from stored scores (ABC::euclidean, src/AbcUtil.cpp:320-324, on Model::scores' output)
// (row i of the set sits at row i + (i >= split ? gap : 0) of T: launch_xb's layout)
__global__ void __launch_bounds__(256) dist_scores_kernel(const double* __restrict__ T, int64_t ldt, int64_t n, int64_t split, int64_t gap, int ncols,
                                                          const double* __restrict__ ref, double* __restrict__ dist) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double* t = T + i + (i >= split ? gap : 0);
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        int c = 0;
        for (; c + 7 < ncols; c += 8) {            // eight independent loads in flight per thread (a pure HBM stream)
            if ((threadIdx.x & 15) == 0 && c + 23 < ncols) {
#pragma unroll
                for (int u = 0; u < 8; u++) asm volatile("prefetch.global.L2 [%0];" ::"l"(t + (int64_t)(c + 16 + u) * ldt));
            }
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = t[(int64_t)(c + u) * ldt];
#pragma unroll
            for (int u = 0; u < 8; u += 4) {
                const double d0 = v[u] - ref[c + u], d1 = v[u + 1] - ref[c + u + 1], d2 = v[u + 2] - ref[c + u + 2], d3 = v[u + 3] - ref[c + u + 3];
                a0 = fma(d0, d0, a0); a1 = fma(d1, d1, a1); a2 = fma(d2, d2, a2); a3 = fma(d3, d3, a3);
            }
        }
        for (; c < ncols; c++) { const double d = t[(int64_t)c * ldt] - ref[c]; a0 = fma(d, d, a0); }
        dist[i] = sqrt((a0 + a1) + (a2 + a3));
    }
}

// S2 + S3 of the ranking as a two-lane pipeline:
//   producer lane   the component loop in blocks of components (a multiple of 32): the one-CTA on-chip loop (pls_defl.cu, state
//                   handed from launch to launch in global memory) or the wide loop's launches (pls_wide.cu, on its SM partition)
//   consumer lane   per finished block: R columns (pls.cpp:412-416), the scores of ALL N rows for those components (T = Zx R,
//                   DMMA), PRESS partial sums + checkpoints of the hold-out rows
// so that the loop (a third of the step at the dengue shape, one SM busy) no longer leaves the other SMs idle. The hold-out
// scores are the bottom rows of T; the final projection (Model::scores on all rows, AbcUtil.cpp:453-454) is its first c* columns.
int rank_fit_holdout_pipelined(abcb200_ctx* ctx, const double* Zx, const double* Zy, int64_t ldz, int64_t N, int64_t n_tr, PlsFactors& fac, double* T_all,
                               int64_t ldt, HoldoutJob* job, const PipePlan& plan) {
    const int K = fac.K, M = fac.M, A = fac.A;
    const int PIPE_BLOCK = plan.block;
    const cudaStream_t lane_small = plan.small, lane_rest = plan.rest;
    const int64_t n_te = N - n_tr, te0 = pad32(n_tr);       // T_all: training rows at [0, n_tr), hold-out rows at [te0, te0 + n_te) of every column
    double* XY = ws_new<double>(ctx, (size_t)K * M);
    double* XX = ws_new<double>(ctx, (size_t)K * K);
    double* U = ws_new<double>(ctx, (size_t)A * A);
    double* state = (plan.kind == 1) ? ws_new<double>(ctx, pls_defl_state_doubles(K, M)) : nullptr;
    if (!XY || !XX || !U || (plan.kind == 1 && !state)) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in the pipelined fit");
    stage_begin(ctx, 1);
    ABC_TRY(launch_gram(ctx, Zx, ldz, K, Zy, ldz, M, n_tr, XX, XY));          // pls.cpp:396, :398
    ABC_TRY(holdout_begin(ctx, Zy + n_tr, ldz, n_te, fac, T_all + te0, ldt, nullptr, job));
    cudaStream_t main_s = ctx->stream;
    WideJob wide;
    if (plan.kind == 2) ABC_TRY(pls_wide_begin(ctx, XX, XY, fac, &wide));
    CUDA_TRY(ctx, cudaEventRecord(ctx->pev[0], main_s));
    CUDA_TRY(ctx, cudaStreamWaitEvent(lane_small, ctx->pev[0], 0));
    CUDA_TRY(ctx, cudaStreamWaitEvent(lane_rest, ctx->pev[0], 0));
    ctx->stat_pls_loop = (plan.kind == 1) ? 1 : 3;
    // Blocks of PIPE_BLOCK components. (Splitting the remainder once more so that the last, unhidden block is narrow was measured and
    // dropped: the scores kernel waits on its operand loads, not on its DMMAs, so a 6-column block costs what a 32-column one does,
    // and the extra launch of the loop plus the extra block of consumer work made C3 0.14 ms and T1M 0.7 ms slower.)
    int bounds[24], nblock = 0;
    bounds[0] = 0;
    for (int c = 0; c < A;) {
        int next = (c + PIPE_BLOCK < A && nblock < 21) ? c + PIPE_BLOCK : A;
        bounds[++nblock] = next; c = next;
    }
    for (int b = 0; b < nblock; b++) {
        const int c0 = bounds[b], c1 = bounds[b + 1];
        {
            StreamScope lane(ctx, lane_small);
            if (b == 0) kernel_begin(ctx, 0);
            if (plan.kind == 1) ABC_TRY(pls_defl_chunk_dev(ctx, XX, XY, fac, c0, c1, state, nullptr));
            else ABC_TRY(pls_wide_block(ctx, &wide, c0, c1));
            if (b == nblock - 1) { kernel_end(ctx, 0); stage_end(ctx, 1); }
            CUDA_TRY(ctx, cudaEventRecord(ctx->pev[1 + b], lane_small));
        }
        {
            StreamScope lane(ctx, lane_rest);
            CUDA_TRY(ctx, cudaStreamWaitEvent(lane_rest, ctx->pev[1 + b], 0));
            if (b == 0) stage_begin(ctx, 2);
            ABC_TRY(pls_ur_block_dev(ctx, fac, U, c0, c1));
            const uint32_t saved = ctx->kernel_timers;
            if (b != 0) ctx->kernel_timers &= ~((1u << 5) | (1u << 4));       // the brackets of xb_kernel<0> / press_chk_kernel time block 0
            kernel_begin(ctx, 5);
            ABC_TRY(launch_xb(ctx, Zx, ldz, N, K, fac.R + (size_t)c0 * K, K, c1 - c0, T_all + (size_t)c0 * ldt, ldt, n_tr, te0 - n_tr));
            kernel_end(ctx, 5);
            ABC_TRY(holdout_press_block(ctx, job, c0, c1));
            ctx->kernel_timers = saved;
            if (b == nblock - 1) {
                ABC_TRY(holdout_press_finalize(ctx, job));
                stage_end(ctx, 2);
                CUDA_TRY(ctx, cudaEventRecord(ctx->pev[1 + nblock], lane_rest));
            }
        }
    }
    CUDA_TRY(ctx, cudaStreamWaitEvent(main_s, ctx->pev[1 + nblock], 0));
    ctx->stat_pipe_block = PIPE_BLOCK;
    return ABCB200_OK;
}

// All pointers are device pointers. order_out: top_n entries (device); dist_out: N (device, nullable).
int rank_core(abcb200_ctx* ctx, const double* met, int64_t ld_met, const double* par, int64_t ld_par, int64_t N, int K, int P,
              const double* target, double f, int method, int64_t top_n, uint64_t* order_out, double* dist_out,
              int* n_comp_used_host, int32_t* n_comp_host, bool simple, const Arrival* arr = nullptr) {
    const int64_t ldz = pad32(N);
    // ---- S1: moments + standardisation (src/AbcUtil.cpp:432-436) ---------------------------------------
    stage_begin(ctx, 0);
    int nchunk = 0;
    double* stats_x = (double*)ws_alloc(ctx, moments_ws_bytes(N, K));
    double* Zx = ws_new<double>(ctx, (size_t)ldz * K);
    double* obs_z = ws_new<double>(ctx, K);
    double* dist = dist_out ? dist_out : ws_new<double>(ctx, N);
    if (!stats_x || !Zx || !obs_z || !dist) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in rank_core");
    double* Zy = nullptr;
    double* stats_y = nullptr;
    if (!simple) {
        stats_y = (double*)ws_alloc(ctx, moments_ws_bytes(N, P));
        Zy = ws_new<double>(ctx, (size_t)ldz * P);
        if (!stats_y || !Zy) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in rank_core");
    }
    if (arr) {
        // the inputs arrive in column blocks: moments and z-scores of a block need nothing but the block (every column is standardised
        // by its own mean and deviation over all rows), so S1 runs under the rest of the copy
        if (!simple) {
            CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, arr->ev_par, 0));
            ABC_TRY(launch_col_stats(ctx, par, ld_par, N, P, stats_y, &nchunk));
            ABC_TRY(launch_zscore(ctx, par, ld_par, N, P, stats_y, nchunk, nullptr, nullptr, Zy, ldz, nullptr, nullptr, nullptr, nullptr));
        }
        for (int b = 0; b < arr->nblk; b++) {
            const int c0 = arr->col0[b], nc = arr->col0[b + 1] - c0;
            CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, arr->ev[b], 0));
            double* st = stats_x + (size_t)c0 * ((N + 2047) / 2048) * 2;        // stats are per column: [col][chunk][2] (moments.cu)
            ABC_TRY(launch_col_stats(ctx, met + (size_t)c0 * ld_met, ld_met, N, nc, st, &nchunk));
            if (b == 0) kernel_begin(ctx, 8);
            ABC_TRY(launch_zscore(ctx, met + (size_t)c0 * ld_met, ld_met, N, nc, st, nchunk, nullptr, nullptr, Zx + (size_t)c0 * ldz, ldz, nullptr, nullptr, target + c0, obs_z + c0));
            if (b == 0) kernel_end(ctx, 8);
        }
    } else {
    ABC_TRY(launch_col_stats(ctx, met, ld_met, N, K, stats_x, &nchunk));
    kernel_begin(ctx, 8);
    ABC_TRY(launch_zscore(ctx, met, ld_met, N, K, stats_x, nchunk, nullptr, nullptr, Zx, ldz, nullptr, nullptr, target, obs_z));
    kernel_end(ctx, 8);
    if (!simple) {
        ABC_TRY(launch_col_stats(ctx, par, ld_par, N, P, stats_y, &nchunk));
        ABC_TRY(launch_zscore(ctx, par, ld_par, N, P, stats_y, nchunk, nullptr, nullptr, Zy, ldz, nullptr, nullptr, nullptr, nullptr));
    }
    }
    stage_end(ctx, 0);

    if (simple) {
        // ---- particle_ranking_simple: distance in z-space (src/AbcUtil.cpp:412-419) ---------------------
        stage_begin(ctx, 4);
        ABC_TRY(launch_euclidean(ctx, Zx, ldz, N, K, obs_z, dist));
        stage_end(ctx, 4);
    } else {
        const int64_t n_tr = (int64_t)std::llround((double)N * f);     // src/AbcUtil.cpp:438
        const int64_t n_te = N - n_tr;
        const int A = K;                                               // 2-argument Model ctor: max_components = X.cols()
        PlsFactors fac;
        fac.K = K; fac.M = P; fac.A = A; fac.method = method; fac.n = n_tr; fac.T = nullptr; fac.ldt = 0;
        fac.W = ws_new<double>(ctx, (size_t)K * A);
        fac.P = ws_new<double>(ctx, (size_t)K * A);
        fac.R = ws_new<double>(ctx, (size_t)K * A);
        fac.Q = ws_new<double>(ctx, (size_t)P * A);
        double* obs_scores = ws_new<double>(ctx, A);
        if (!fac.W || !fac.P || !fac.R || !fac.Q || !obs_scores) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in rank_core");
        int32_t ncomp_local[128];
        int32_t* ncomp = n_comp_host ? n_comp_host : ncomp_local;
        ctx->stat_pipe_block = 0;
        const PipePlan plan = rank_pipe_plan(ctx, K, P, method, n_te);
        if (plan.kind != 0) {
            // ---- S2 + S3 as a two-lane pipeline, S4 after it; the projection is already there (the first c* columns of T) ----
            const int64_t ldt = pad32(n_tr) + pad32(n_te);
            double* T_all = ws_new<double>(ctx, (size_t)ldt * A);
            if (!T_all) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in rank_core");
            HoldoutJob job;
            ABC_TRY(rank_fit_holdout_pipelined(ctx, Zx, Zy, ldz, N, n_tr, fac, T_all, ldt, &job, plan));
            ABC_TRY(holdout_select_finish(ctx, &job, 0.1, ncomp));
            int used = 0;
            for (int y = 0; y < P; y++) used = ncomp[y] > used ? ncomp[y] : used;
            if (n_comp_used_host) *n_comp_used_host = used;
            stage_begin(ctx, 4);
            ABC_TRY(launch_vec_times_mat(ctx, obs_z, K, fac.R, K, used, obs_scores));
            const int dgrid = (int)std::max<int64_t>(1, std::min<int64_t>((N + 255) / 256, (int64_t)8 * ctx->sm_count));
            LAUNCH(ctx, dist_scores_kernel, dgrid, 256, 0, T_all, ldt, N, n_tr, pad32(n_tr) - n_tr, used, obs_scores, dist);
            stage_end(ctx, 4);
        } else {
        // ---- S2: PLS fit on the training rows (src/AbcUtil.cpp:443) --------------------------------------
        stage_begin(ctx, 1);
        ABC_TRY(pls_fit_dev(ctx, Zx, ldz, Zy, ldz, fac));
        stage_end(ctx, 1);
        // ---- S3 + S4: hold-out validation and component selection (src/AbcUtil.cpp:446-449) ---------------
        ABC_TRY(holdout_select_dev(ctx, Zx + n_tr, ldz, Zy + n_tr, ldz, n_te, fac, 0.1, nullptr, ncomp));
        int used = 0;
        for (int y = 0; y < P; y++) used = ncomp[y] > used ? ncomp[y] : used;
        if (n_comp_used_host) *n_comp_used_host = used;
        // ---- S5: projection of every particle fused with the distance (src/AbcUtil.cpp:453-455) -----------
        stage_begin(ctx, 4);
        ABC_TRY(launch_vec_times_mat(ctx, obs_z, K, fac.R, K, used, obs_scores));
        kernel_begin(ctx, 6);
        ABC_TRY(launch_project_dist(ctx, Zx, ldz, N, K, fac.R, K, used, obs_scores, dist));
        kernel_end(ctx, 6);
        stage_end(ctx, 4);
        }
    }
    // ---- S6: ordering (src/AbcUtil.cpp:457, AbcSmc.cpp:645-646) ---------------------------------------------
    stage_begin(ctx, 5);
    const int rc = order_dev(ctx, dist, N, top_n, order_out);
    stage_end(ctx, 5);
    return rc;
}

}  // namespace

// H2D of one set's inputs on the copy stream: the target and the parameters first, then the metrics in up to 8 column blocks, an event
// after each; rank_core starts the moments / z-scores of a block as soon as it has arrived (everything after S1 needs ALL columns'
// moments over ALL rows, so that is as far as the overlap can go: DESIGN.md 8). P == 0: no parameter matrix (SIMPLE filter).
int stage_inputs(abcb200_ctx* ctx, double* d_met, double* d_par, double* d_target, int64_t ldd, const double* met, int64_t ld_met, const double* par,
                 int64_t ld_par, const double* target, int64_t N, int K, int P, Arrival* arr) {
    StreamScope cs(ctx, ctx->copy_stream);
    stage_begin(ctx, 8);
    CUDA_TRY(ctx, cudaMemcpyAsync(d_target, target, sizeof(double) * K, cudaMemcpyHostToDevice, ctx->stream));
    if (P > 0) ABC_TRY(h2d_matrix(ctx, d_par, ldd, par, ld_par, N, P));
    CUDA_TRY(ctx, cudaEventRecord(ctx->cev[10], ctx->stream));
    arr->ev_par = ctx->cev[10];
    const int nblk = (int)std::max<int64_t>(1, std::min<int64_t>(8, std::min<int64_t>(K, ((int64_t)N * K * 8) >> 23)));     // >= 8 MB per block
    arr->nblk = nblk;
    for (int b = 0; b <= nblk; b++) arr->col0[b] = (int)((int64_t)K * b / nblk);
    for (int b = 0; b < nblk; b++) {
        const int c0 = arr->col0[b], nc = arr->col0[b + 1] - c0;
        ABC_TRY(h2d_matrix(ctx, d_met + (size_t)c0 * ldd, ldd, met + (size_t)c0 * ld_met, ld_met, N, nc));
        CUDA_TRY(ctx, cudaEventRecord(ctx->cev[b], ctx->stream));
        arr->ev[b] = ctx->cev[b];
    }
    stage_end(ctx, 8);
    return ABCB200_OK;
}

// ---- exact distance ties as the reference leaves them (abcb200_set_tie_order 1) ------------------------------------------------------
// PLS::ordered (lib/PLS/include/PLS/pls.h:58-69; the same construction as lib/ranker.h:47-53) is std::iota + an index std::sort with a
// strict < on the distances. std::sort is not stable: where EQUAL distances land is decided by libstdc++'s introsort run over all N
// indices (partition swaps, the final insertion sort), not by the algorithm this library implements, and no local rule reproduces it.
// The device order breaks ties by ascending particle index. Everything else is unambiguous: if no two of the top_n returned distances
// are equal and no distance beyond the cut equals the last one, the first top_n entries of ANY ascending order — std::sort's included —
// are the device order. Only when exact ties do reach the output (duplicated particles, integer-valued metrics such as the dice game
// of examples/) is the reference's statement executed here, on the GPU-computed distances, by the same libstdc++ this file is built with.
bool tie_order_stdsort(const double* dist, int64_t N, int64_t top_n, uint64_t* order) {
    bool tie = false;
    for (int64_t i = 1; i < top_n && !tie; i++) tie = dist[order[i]] == dist[order[i - 1]];
    if (!tie && top_n < N) {
        const double last = dist[order[top_n - 1]];
        int64_t not_after = 0;
        for (int64_t i = 0; i < N; i++) not_after += dist[i] <= last;
        tie = not_after != top_n;                    // a particle beyond the cut is as close as the last one kept
    }
    if (!tie) return false;
    std::vector<size_t> idx((size_t)N);
    std::iota(idx.begin(), idx.end(), (size_t)0);
    std::sort(idx.begin(), idx.end(), [dist](const size_t& lhs, const size_t& rhs) { return dist[lhs] < dist[rhs]; });
    for (int64_t i = 0; i < top_n; i++) order[i] = (uint64_t)idx[(size_t)i];
    return true;
}

extern "C" int abcb200_tie_order_stdsort(const double* dist, int64_t N, int64_t top_n, uint64_t* order) {
    if (!dist || !order || N < 1 || top_n < 1 || top_n > N) return ABCB200_EINVAL;
    for (int64_t i = 0; i < top_n; i++) if (order[i] >= (uint64_t)N) return ABCB200_EINVAL;
    for (int64_t i = 0; i < N; i++) if (dist[i] != dist[i]) return ABCB200_ENAN;      // std::sort's comparator would be inconsistent (as abcb200_ordered)
    return tie_order_stdsort(dist, N, top_n, order) ? 1 : 0;
}

int tie_order_stdsort_device(abcb200_ctx* ctx, const double* d_dist, int64_t N, int64_t top_n, uint64_t* d_order) {
    ABC_TRY(hpin_reserve(ctx, sizeof(double) * (size_t)N + sizeof(uint64_t) * (size_t)top_n + 64));     // pinned: full-rate copies, no page faults
    double* dist = (double*)ctx->hpin;
    uint64_t* order = (uint64_t*)(dist + N);
    CUDA_TRY(ctx, cudaMemcpyAsync(dist, d_dist, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(order, d_order, sizeof(uint64_t) * (size_t)top_n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (!tie_order_stdsort(dist, N, top_n, order)) return ABCB200_OK;
    ctx->stat_tie_resorts++;
    CUDA_TRY(ctx, cudaMemcpyAsync(d_order, order, sizeof(uint64_t) * (size_t)top_n, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));      // the pinned scratch is reused by the next call
    return ABCB200_OK;
}

namespace {

int rank_check(abcb200_ctx* ctx, int64_t N, int K, int P, double f, int method, bool simple) {
    if (N < 2 || K < 1) ABC_FAIL(ctx, ABCB200_EINVAL, "rank: need N >= 2 and K >= 1 (N=%lld K=%d)", (long long)N, K);
    if (simple) return ABCB200_OK;
    if (P < 1 || P > 128) ABC_FAIL(ctx, ABCB200_EINVAL, "rank_pls: number of parameters P=%d outside [1,128]", P);
    if (!(f > 0.0 && f <= 1.0)) ABC_FAIL(ctx, ABCB200_EINVAL, "rank_pls: training_fraction %g outside (0,1] (AbcUtil.cpp:428)", f);
    if (method < ABCB200_KERNEL_TYPE1 || method > ABCB200_KERNEL_TYPE1_STREAM) ABC_FAIL(ctx, ABCB200_EINVAL, "rank_pls: unknown method %d", method);
    const int64_t n_tr = (int64_t)std::llround((double)N * f);
    if (n_tr < K) ABC_FAIL(ctx, ABCB200_EINVAL, "rank_pls: %lld training rows < K=%d components (tt ~ 0; undefined in the reference)", (long long)n_tr, K);
    return ABCB200_OK;
}

int rank_host(abcb200_ctx* ctx, const double* met, int64_t ld_met, const double* par, int64_t ld_par, int64_t N, int K, int P,
              const double* target, double f, int method, int64_t top_n, uint64_t* order_out, double* dist_out,
              int* n_comp_used_out, int32_t* n_comp_out, bool simple) {
    ABC_TRY(check_ctx(ctx));
    if (!met || !target || !order_out || (!simple && !par)) ABC_FAIL(ctx, ABCB200_EINVAL, "rank: null argument");
    if (ld_met < N || (!simple && ld_par < N)) ABC_FAIL(ctx, ABCB200_EINVAL, "rank: leading dimension < N");
    ABC_TRY(rank_check(ctx, N, K, P, f, method, simple));
    if (top_n <= 0 || top_n > N) top_n = N;
    if (ctx->tie_order == 1 && !dist_out) ABC_TRY(hpin_reserve(ctx, sizeof(double) * (size_t)N + 64));
    const int64_t ldd = pad32(N);
    size_t need = rank_core_ws_bytes(ctx, N, K, P, f, method, simple);
    need += align_up((size_t)ldd * K * 8, 256) + (simple ? 0 : align_up((size_t)ldd * P * 8, 256)) + align_up((size_t)K * 8, 256) +
            align_up((size_t)N * 8, 256) + align_up((size_t)N * 8, 256) + 4096;
    ABC_TRY(ws_reserve(ctx, need));
    double* d_met = ws_new<double>(ctx, (size_t)ldd * K);
    double* d_par = simple ? nullptr : ws_new<double>(ctx, (size_t)ldd * P);
    double* d_target = ws_new<double>(ctx, K);
    double* d_dist = ws_new<double>(ctx, N);
    uint64_t* d_order = ws_new<uint64_t>(ctx, N);
    if (!d_met || (!simple && !d_par) || !d_target || !d_dist || !d_order) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in rank");
    Arrival arr;
    ABC_TRY(stage_inputs(ctx, d_met, d_par, d_target, ldd, met, ld_met, par, ld_par, target, N, K, simple ? 0 : P, &arr));
    ABC_TRY(rank_core(ctx, d_met, ldd, d_par, ldd, N, K, P, d_target, f, method, top_n, d_order, d_dist, n_comp_used_out, n_comp_out, simple, &arr));
    if (ctx->tie_order == 1 && !dist_out) dist_out = (double*)ctx->hpin;      // reserved above: the pinned scratch is free once rank_core has returned
    stage_begin(ctx, 9);
    ABC_TRY(d2h(ctx, order_out, d_order, sizeof(uint64_t) * (size_t)top_n));
    if (dist_out) ABC_TRY(d2h(ctx, dist_out, d_dist, sizeof(double) * (size_t)N));
    stage_end(ctx, 9);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->tie_order == 1 && tie_order_stdsort(dist_out, N, top_n, order_out)) ctx->stat_tie_resorts++;
    return ABCB200_OK;
}

int rank_dev(abcb200_ctx* ctx, const double* met, int64_t ld_met, const double* par, int64_t ld_par, int64_t N, int K, int P,
             const double* target, double f, int method, int64_t top_n, uint64_t* order_out, double* dist_out,
             int* n_comp_used_out, int32_t* n_comp_out, bool simple) {
    ABC_TRY(check_ctx(ctx));
    if (!met || !target || !order_out || (!simple && !par)) ABC_FAIL(ctx, ABCB200_EINVAL, "rank: null argument");
    if (ld_met < N || (!simple && ld_par < N)) ABC_FAIL(ctx, ABCB200_EINVAL, "rank: leading dimension < N");
    ABC_TRY(rank_check(ctx, N, K, P, f, method, simple));
    if (top_n <= 0 || top_n > N) top_n = N;
    ABC_TRY(ws_reserve(ctx, rank_core_ws_bytes(ctx, N, K, P, f, method, simple) + align_up((size_t)N * 8, 256)));
    if (ctx->tie_order == 1 && !dist_out) {
        dist_out = ws_new<double>(ctx, (size_t)N);
        if (!dist_out) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in rank");
    }
    ABC_TRY(rank_core(ctx, met, ld_met, par, ld_par, N, K, P, target, f, method, top_n, order_out, dist_out, n_comp_used_out, n_comp_out, simple));
    if (ctx->tie_order == 1) ABC_TRY(tie_order_stdsort_device(ctx, dist_out, N, top_n, order_out));
    return ABCB200_OK;
}

}  // namespace

// the ranking on device-resident inputs, for chain.cu (everything stays in the caller's workspace reservation)
size_t rank_ws_bytes(const abcb200_ctx* ctx, int64_t N, int K, int P, double f, int method, bool simple) { return rank_core_ws_bytes(ctx, N, K, P, f, method, simple); }
int rank_shape_check(abcb200_ctx* ctx, int64_t N, int K, int P, double f, int method, bool simple) { return rank_check(ctx, N, K, P, f, method, simple); }
int rank_on_device(abcb200_ctx* ctx, const double* met, int64_t ld_met, const double* par, int64_t ld_par, int64_t N, int K, int P, const double* target,
                   double f, int method, int64_t top_n, uint64_t* order_out, double* dist_out, int* n_comp_used_host, int32_t* n_comp_host, bool simple,
                   const Arrival* arr) {
    return rank_core(ctx, met, ld_met, par, ld_par, N, K, P, target, f, method, top_n, order_out, dist_out, n_comp_used_host, n_comp_host, simple, arr);
}

namespace {

// elementwise / small reductions for the Model API
__global__ void residual_kernel(const double* __restrict__ Y, int64_t ldy, const double* __restrict__ F, int64_t ldf, int64_t n, int M,
                                double* __restrict__ out, int64_t ldo) {
    const int m = blockIdx.y;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[(int64_t)m * ldo + i] = Y[(int64_t)m * ldy + i] - F[(int64_t)m * ldf + i];
}
__global__ void colsumsq_kernel(const double* __restrict__ E, int64_t ld, int64_t n, double* __restrict__ out) {
    __shared__ double red[32];
    const int m = blockIdx.x;
    double s = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { const double e = E[(int64_t)m * ld + i]; s = fma(e, e, s); }
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[m] = s;
}

}  // namespace

// =================================================================================================================
extern "C" int abcb200_rank_pls(abcb200_ctx* ctx, const double* met, int64_t ld_met, const double* par, int64_t ld_par, int64_t N,
                                int K, int P, const double* target, double training_fraction, int method, int64_t top_n,
                                uint64_t* order_out, double* dist_out, int* n_comp_used_out, int32_t* n_comp_per_y_out) {
    return rank_host(ctx, met, ld_met, par, ld_par, N, K, P, target, training_fraction, method, top_n, order_out, dist_out,
                     n_comp_used_out, n_comp_per_y_out, false);
}
extern "C" int abcb200_rank_pls_dev(abcb200_ctx* ctx, const double* met, int64_t ld_met, const double* par, int64_t ld_par, int64_t N,
                                    int K, int P, const double* target, double training_fraction, int method, int64_t top_n,
                                    uint64_t* order_out, double* dist_out, int* n_comp_used_out, int32_t* n_comp_per_y_out) {
    return rank_dev(ctx, met, ld_met, par, ld_par, N, K, P, target, training_fraction, method, top_n, order_out, dist_out,
                    n_comp_used_out, n_comp_per_y_out, false);
}
extern "C" int abcb200_rank_simple(abcb200_ctx* ctx, const double* met, int64_t ld_met, int64_t N, int K, const double* target,
                                   int64_t top_n, uint64_t* order_out, double* dist_out) {
    return rank_host(ctx, met, ld_met, nullptr, 0, N, K, 0, target, 0.5, 0, top_n, order_out, dist_out, nullptr, nullptr, true);
}
extern "C" int abcb200_rank_simple_dev(abcb200_ctx* ctx, const double* met, int64_t ld_met, int64_t N, int K, const double* target,
                                       int64_t top_n, uint64_t* order_out, double* dist_out) {
    return rank_dev(ctx, met, ld_met, nullptr, 0, N, K, 0, target, 0.5, 0, top_n, order_out, dist_out, nullptr, nullptr, true);
}

// ---- doubled variance ---------------------------------------------------------------------------------------
static int dv_core(abcb200_ctx* ctx, const double* params, int64_t ld, int64_t n, int P, double* dv_out) {
    stage_begin(ctx, 6);
    int nchunk = 0;
    double* stats = (double*)ws_alloc(ctx, moments_ws_bytes(n, P));
    if (!stats) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in doubled_variance");
    ABC_TRY(launch_col_stats(ctx, params, ld, n, P, stats, &nchunk));
    ABC_TRY(launch_col_finalize(ctx, stats, nchunk, n, P, nullptr, nullptr, 2.0, dv_out));   // 2 * sample variance (AbcUtil.cpp:534)
    stage_end(ctx, 6);
    return ABCB200_OK;
}
extern "C" int abcb200_doubled_variance_dev(abcb200_ctx* ctx, const double* params, int64_t ld, int64_t n, int P, double* dv_out) {
    ABC_TRY(check_ctx(ctx));
    if (!params || !dv_out || n < 1 || P < 1 || ld < n) ABC_FAIL(ctx, ABCB200_EINVAL, "doubled_variance: bad argument");
    ABC_TRY(ws_reserve(ctx, moments_ws_bytes(n, P) + 1024));
    return dv_core(ctx, params, ld, n, P, dv_out);
}
extern "C" int abcb200_doubled_variance_gather_dev(abcb200_ctx* ctx, const double* params, int64_t ld, const uint64_t* idx, int64_t n,
                                                   int P, double* gathered_out, double* dv_out) {
    ABC_TRY(check_ctx(ctx));
    if (!params || !idx || !dv_out || n < 1 || P < 1 || ld < 1) ABC_FAIL(ctx, ABCB200_EINVAL, "doubled_variance_gather: bad argument");
    ABC_TRY(ws_reserve(ctx, moments_ws_bytes(n, P) + align_up((size_t)n * P * 8, 256) + 2048));
    double* g = gathered_out ? gathered_out : ws_new<double>(ctx, (size_t)n * P);
    if (!g) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in doubled_variance_gather");
    ABC_TRY(launch_gather_rows(ctx, params, ld, idx, n, P, g, n));
    return dv_core(ctx, g, n, n, P, dv_out);
}
extern "C" int abcb200_doubled_variance(abcb200_ctx* ctx, const double* params, int64_t ld, int64_t n, int P, double* dv_out) {
    ABC_TRY(check_ctx(ctx));
    if (!params || !dv_out || n < 1 || P < 1 || ld < n) ABC_FAIL(ctx, ABCB200_EINVAL, "doubled_variance: bad argument");
    const int64_t ldd = pad32(n);
    ABC_TRY(ws_reserve(ctx, moments_ws_bytes(n, P) + align_up((size_t)ldd * P * 8, 256) + align_up((size_t)P * 8, 256) + 2048));
    double* d = ws_new<double>(ctx, (size_t)ldd * P);
    double* d_out = ws_new<double>(ctx, P);
    if (!d || !d_out) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in doubled_variance");
    stage_begin(ctx, 8);
    ABC_TRY(h2d_matrix(ctx, d, ldd, params, ld, n, P));
    stage_end(ctx, 8);
    ABC_TRY(dv_core(ctx, d, ldd, n, P, d_out));
    stage_begin(ctx, 9);
    ABC_TRY(d2h(ctx, dv_out, d_out, sizeof(double) * P));
    stage_end(ctx, 9);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ABCB200_OK;
}

// ---- next-set proposal sampling -----------------------------------------------------------
static int sample_check(abcb200_ctx* ctx, int64_t num_samples, int64_t ld, int64_t n_pp, int P, int max_attempts, int64_t ld_out) {
    if (num_samples < 1 || n_pp < 1 || P < 1 || ld < n_pp || ld_out < num_samples || max_attempts < 1)
        ABC_FAIL(ctx, ABCB200_EINVAL, "sample_predictive_priors: bad shape (num_samples=%lld n_pp=%lld P=%d)", (long long)num_samples, (long long)n_pp, P);
    return ABCB200_OK;
}
extern "C" int abcb200_sample_predictive_priors_dev(abcb200_ctx* ctx, uint64_t seed, int64_t num_samples, const double* weights,
                                                    const double* theta, int64_t ld, int64_t n_pp, int P, const double* dv, const double* lo,
                                                    const double* hi, const int32_t* integral, const double* prior_mean, int max_attempts,
                                                    double* out, int64_t ld_out, uint64_t* parent_out, uint64_t* fallbacks_out) {
    ABC_TRY(check_ctx(ctx));
    if (!weights || !theta || !dv || !lo || !hi || !prior_mean || !out) ABC_FAIL(ctx, ABCB200_EINVAL, "sample_predictive_priors: null argument");
    ABC_TRY(sample_check(ctx, num_samples, ld, n_pp, P, max_attempts, ld_out));
    ABC_TRY(ws_reserve(ctx, sample_ws_bytes(n_pp) + 1024));
    if (fallbacks_out) CUDA_TRY(ctx, cudaMemsetAsync(fallbacks_out, 0, sizeof(uint64_t), ctx->stream));
    return sample_predictive_priors_core(ctx, seed, num_samples, weights, theta, ld, n_pp, P, dv, lo, hi, integral, prior_mean, max_attempts, out,
                                         ld_out, parent_out, (unsigned long long*)fallbacks_out);
}
extern "C" int abcb200_sample_predictive_priors(abcb200_ctx* ctx, uint64_t seed, int64_t num_samples, const double* weights, const double* theta,
                                                int64_t ld, int64_t n_pp, int P, const double* dv, const double* lo, const double* hi,
                                                const int32_t* integral, const double* prior_mean, int max_attempts, double* out,
                                                int64_t ld_out, uint64_t* parent_out, uint64_t* fallbacks_out) {
    ABC_TRY(check_ctx(ctx));
    if (!weights || !theta || !dv || !lo || !hi || !prior_mean || !out) ABC_FAIL(ctx, ABCB200_EINVAL, "sample_predictive_priors: null argument");
    ABC_TRY(sample_check(ctx, num_samples, ld, n_pp, P, max_attempts, ld_out));
    double total = 0.0;                       // gsl_ran_discrete_preproc aborts on a negative weight or an all-zero table
    for (int64_t j = 0; j < n_pp; j++) {
        if (!(weights[j] >= 0.0) || !std::isfinite(weights[j])) ABC_FAIL(ctx, ABCB200_EINVAL, "sample_predictive_priors: weight %lld is negative or not finite", (long long)j);
        total += weights[j];
    }
    if (!(total > 0.0)) ABC_FAIL(ctx, ABCB200_EINVAL, "sample_predictive_priors: all weights are zero");
    for (int p = 0; p < P; p++)
        if (!(dv[p] >= 0.0)) ABC_FAIL(ctx, ABCB200_EINVAL, "sample_predictive_priors: doubled variance %d is negative or NaN", p);
    const int64_t ldt = pad32(n_pp), ldo = pad32(num_samples);
    size_t need = sample_ws_bytes(n_pp) + align_up((size_t)ldt * P * 8, 256) + align_up((size_t)ldo * P * 8, 256) + align_up((size_t)n_pp * 8, 256) +
                  align_up((size_t)num_samples * 8, 256) + 6 * align_up((size_t)P * 8, 256) + 4096;
    ABC_TRY(ws_reserve(ctx, need));
    double* d_theta = ws_new<double>(ctx, (size_t)ldt * P);
    double* d_out = ws_new<double>(ctx, (size_t)ldo * P);
    double* d_w = ws_new<double>(ctx, n_pp);
    uint64_t* d_parent = parent_out ? ws_new<uint64_t>(ctx, num_samples) : nullptr;
    double* d_dv = ws_new<double>(ctx, P);
    double* d_lo = ws_new<double>(ctx, P);
    double* d_hi = ws_new<double>(ctx, P);
    double* d_mean = ws_new<double>(ctx, P);
    int32_t* d_int = integral ? ws_new<int32_t>(ctx, P) : nullptr;
    unsigned long long* d_fb = ws_new<unsigned long long>(ctx, 1);
    if (!d_theta || !d_out || !d_w || !d_dv || !d_lo || !d_hi || !d_mean || !d_fb) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in sample_predictive_priors");
    stage_begin(ctx, 8);
    ABC_TRY(h2d_matrix(ctx, d_theta, ldt, theta, ld, n_pp, P));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_w, weights, sizeof(double) * n_pp, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_dv, dv, sizeof(double) * P, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_lo, lo, sizeof(double) * P, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_hi, hi, sizeof(double) * P, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_mean, prior_mean, sizeof(double) * P, cudaMemcpyHostToDevice, ctx->stream));
    if (integral) CUDA_TRY(ctx, cudaMemcpyAsync(d_int, integral, sizeof(int32_t) * P, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(d_fb, 0, sizeof(unsigned long long), ctx->stream));
    stage_end(ctx, 8);
    ABC_TRY(sample_predictive_priors_core(ctx, seed, num_samples, d_w, d_theta, ldt, n_pp, P, d_dv, d_lo, d_hi, d_int, d_mean, max_attempts, d_out, ldo,
                                          d_parent, d_fb));
    stage_begin(ctx, 9);
    CUDA_TRY(ctx, cudaMemcpy2DAsync(out, (size_t)ld_out * 8, d_out, (size_t)ldo * 8, (size_t)num_samples * 8, (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
    if (parent_out) ABC_TRY(d2h(ctx, parent_out, d_parent, sizeof(uint64_t) * (size_t)num_samples));
    unsigned long long fb = 0;
    ABC_TRY(d2h(ctx, &fb, d_fb, sizeof(fb)));
    stage_end(ctx, 9);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (fallbacks_out) *fallbacks_out = (uint64_t)fb;
    return ABCB200_OK;
}

// ABC::setup_mvn_sampler (src/AbcUtil.cpp:462-488): L_out = P x P column-major lower Cholesky factor (host)
extern "C" int abcb200_setup_mvn_sampler(abcb200_ctx* ctx, const double* theta, int64_t ld, int64_t n_pp, int P, double* L_out) {
    ABC_TRY(check_ctx(ctx));
    if (!theta || !L_out || n_pp < 2 || P < 1 || P > 128 || ld < n_pp) ABC_FAIL(ctx, ABCB200_EINVAL, "setup_mvn_sampler: bad argument (n_pp=%lld P=%d)", (long long)n_pp, P);
    const int64_t ldt = pad32(n_pp);
    ABC_TRY(ws_reserve(ctx, mvn_setup_ws_bytes(n_pp, P) + align_up((size_t)ldt * P * 8, 256) + align_up((size_t)P * P * 8, 256) + 4096));
    double* d_theta = ws_new<double>(ctx, (size_t)ldt * P);
    double* d_L = ws_new<double>(ctx, (size_t)P * P);
    int* d_flag = ws_new<int>(ctx, 1);
    if (!d_theta || !d_L || !d_flag) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in setup_mvn_sampler");
    ABC_TRY(h2d_matrix(ctx, d_theta, ldt, theta, ld, n_pp, P));
    CUDA_TRY(ctx, cudaMemsetAsync(d_flag, 0, sizeof(int), ctx->stream));
    ABC_TRY(setup_mvn_sampler_core(ctx, d_theta, ldt, n_pp, P, d_L, d_flag));
    int flag = 0;
    ABC_TRY(d2h(ctx, L_out, d_L, sizeof(double) * (size_t)P * P));
    ABC_TRY(d2h(ctx, &flag, d_flag, sizeof(int)));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (flag) ABC_FAIL(ctx, ABCB200_EINVAL, "setup_mvn_sampler: the doubled-diagonal covariance is not positive definite (gsl_linalg_cholesky_decomp1: GSL_EDOM)");
    return ABCB200_OK;
}
// ABC::sample_mvn_predictive_priors (src/AbcUtil.cpp:392-404) with ABC::gsl_ran_trunc_mv_normal (:123-144); host buffers
extern "C" int abcb200_sample_mvn_predictive_priors(abcb200_ctx* ctx, uint64_t seed, int64_t num_samples, const double* weights, const double* theta,
                                                    int64_t ld, int64_t n_pp, int P, const double* L, const double* lo, const double* hi,
                                                    const int32_t* integral, int max_attempts, double* out, int64_t ld_out, uint64_t* parent_out,
                                                    uint64_t* failures_out) {
    ABC_TRY(check_ctx(ctx));
    if (!weights || !theta || !L || !lo || !hi || !out) ABC_FAIL(ctx, ABCB200_EINVAL, "sample_mvn_predictive_priors: null argument");
    ABC_TRY(sample_check(ctx, num_samples, ld, n_pp, P, max_attempts, ld_out));
    if (P > 128) ABC_FAIL(ctx, ABCB200_EINVAL, "sample_mvn_predictive_priors: P=%d exceeds 128", P);
    double total = 0.0;
    for (int64_t j = 0; j < n_pp; j++) {
        if (!(weights[j] >= 0.0) || !std::isfinite(weights[j])) ABC_FAIL(ctx, ABCB200_EINVAL, "sample_mvn_predictive_priors: weight %lld is negative or not finite", (long long)j);
        total += weights[j];
    }
    if (!(total > 0.0)) ABC_FAIL(ctx, ABCB200_EINVAL, "sample_mvn_predictive_priors: all weights are zero");
    const int64_t ldt = pad32(n_pp), ldo = pad32(num_samples);
    size_t need = sample_ws_bytes(n_pp) + align_up((size_t)ldt * P * 8, 256) + align_up((size_t)ldo * P * 8, 256) + align_up((size_t)n_pp * 8, 256) +
                  align_up((size_t)num_samples * 8, 256) + align_up((size_t)P * P * 8, 256) + 4 * align_up((size_t)P * 8, 256) + 4096;
    ABC_TRY(ws_reserve(ctx, need));
    double* d_theta = ws_new<double>(ctx, (size_t)ldt * P);
    double* d_out = ws_new<double>(ctx, (size_t)ldo * P);
    double* d_w = ws_new<double>(ctx, n_pp);
    uint64_t* d_parent = parent_out ? ws_new<uint64_t>(ctx, num_samples) : nullptr;
    double* d_L = ws_new<double>(ctx, (size_t)P * P);
    double* d_lo = ws_new<double>(ctx, P);
    double* d_hi = ws_new<double>(ctx, P);
    int32_t* d_int = integral ? ws_new<int32_t>(ctx, P) : nullptr;
    unsigned long long* d_fail = ws_new<unsigned long long>(ctx, 1);
    if (!d_theta || !d_out || !d_w || !d_L || !d_lo || !d_hi || !d_fail) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in sample_mvn_predictive_priors");
    stage_begin(ctx, 8);
    ABC_TRY(h2d_matrix(ctx, d_theta, ldt, theta, ld, n_pp, P));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_w, weights, sizeof(double) * n_pp, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_L, L, sizeof(double) * P * P, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_lo, lo, sizeof(double) * P, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_hi, hi, sizeof(double) * P, cudaMemcpyHostToDevice, ctx->stream));
    if (integral) CUDA_TRY(ctx, cudaMemcpyAsync(d_int, integral, sizeof(int32_t) * P, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(d_fail, 0, sizeof(unsigned long long), ctx->stream));
    stage_end(ctx, 8);
    ABC_TRY(sample_mvn_predictive_priors_core(ctx, seed, num_samples, d_w, d_theta, ldt, n_pp, P, d_L, d_lo, d_hi, d_int, max_attempts, d_out, ldo, d_parent, d_fail));
    stage_begin(ctx, 9);
    CUDA_TRY(ctx, cudaMemcpy2DAsync(out, (size_t)ld_out * 8, d_out, (size_t)ldo * 8, (size_t)num_samples * 8, (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
    if (parent_out) ABC_TRY(d2h(ctx, parent_out, d_parent, sizeof(uint64_t) * (size_t)num_samples));
    unsigned long long fl = 0;
    ABC_TRY(d2h(ctx, &fl, d_fail, sizeof(fl)));
    stage_end(ctx, 9);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (failures_out) *failures_out = (uint64_t)fl;
    return ABCB200_OK;
}

// ---- weights -------------------------------------------------------------------------------------------------
extern "C" int abcb200_weights_set0(abcb200_ctx* ctx, int64_t n, double* w_out) {
    ABC_TRY(check_ctx(ctx));
    if (n < 1 || !w_out) ABC_FAIL(ctx, ABCB200_EINVAL, "weights_set0: bad argument");
    ABC_TRY(ws_reserve(ctx, align_up((size_t)n * 8, 256) + 1024));
    double* d = ws_new<double>(ctx, n);
    ABC_TRY(launch_fill(ctx, d, n, 1.0 / (double)n));          // src/AbcUtil.cpp:543-544
    ABC_TRY(d2h(ctx, w_out, d, sizeof(double) * (size_t)n));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ABCB200_OK;
}

extern "C" int abcb200_weights_unnorm_dev(abcb200_ctx* ctx, const double* numer, const double* theta_new, int64_t ld_new, int64_t n_rows,
                                          const double* theta_old, int64_t ld_old, int64_t N_old, const double* w_old,
                                          const double* dv_old, int P, int algo, double* w_out, double* sumsq_out) {
    ABC_TRY(check_ctx(ctx));
    if (!theta_new || !theta_old || !w_old || !dv_old || !w_out || !sumsq_out) ABC_FAIL(ctx, ABCB200_EINVAL, "weights: null argument");
    if (n_rows < 0 || N_old < 1 || P < 1 || ld_new < n_rows || ld_old < N_old) ABC_FAIL(ctx, ABCB200_EINVAL, "weights: bad shape");
    ABC_TRY(ws_reserve(ctx, weights_ws_bytes(ctx, n_rows, N_old, P)));
    stage_begin(ctx, 7);
    const int rc = weights_unnorm_dev(ctx, numer, theta_new, ld_new, n_rows, theta_old, ld_old, N_old, w_old, dv_old, P, algo, w_out, sumsq_out);
    stage_end(ctx, 7);
    return rc;
}

extern "C" int abcb200_scale_weights_dev(abcb200_ctx* ctx, double* w, int64_t n, const double* sumsq) {
    ABC_TRY(check_ctx(ctx));
    if (!w || !sumsq || n < 0) ABC_FAIL(ctx, ABCB200_EINVAL, "scale_weights: bad argument");
    if (n == 0) return ABCB200_OK;
    return launch_scale_weights(ctx, w, n, sumsq);
}

extern "C" int abcb200_weights_dev(abcb200_ctx* ctx, const double* numer, const double* theta_new, int64_t ld_new, int64_t N_new,
                                   const double* theta_old, int64_t ld_old, int64_t N_old, const double* w_old, const double* dv_old,
                                   int P, int algo, double* w_out) {
    ABC_TRY(check_ctx(ctx));
    if (!theta_new || !theta_old || !w_old || !dv_old || !w_out) ABC_FAIL(ctx, ABCB200_EINVAL, "weights: null argument");
    if (N_new < 1 || N_old < 1 || P < 1 || ld_new < N_new || ld_old < N_old) ABC_FAIL(ctx, ABCB200_EINVAL, "weights: bad shape");
    ABC_TRY(ws_reserve(ctx, weights_ws_bytes(ctx, N_new, N_old, P) + 1024));
    double* ss = ws_new<double>(ctx, 1);
    stage_begin(ctx, 7);
    ABC_TRY(weights_unnorm_dev(ctx, numer, theta_new, ld_new, N_new, theta_old, ld_old, N_old, w_old, dv_old, P, algo, w_out, ss));
    ABC_TRY(launch_scale_weights(ctx, w_out, N_new, ss));
    stage_end(ctx, 7);
    return ABCB200_OK;
}

extern "C" int abcb200_weights(abcb200_ctx* ctx, const double* numer, const double* theta_new, int64_t ld_new, int64_t N_new,
                               const double* theta_old, int64_t ld_old, int64_t N_old, const double* w_old, const double* dv_old, int P,
                               int algo, double* w_out) {
    ABC_TRY(check_ctx(ctx));
    if (!theta_new || !theta_old || !w_old || !dv_old || !w_out) ABC_FAIL(ctx, ABCB200_EINVAL, "weights: null argument");
    if (N_new < 1 || N_old < 1 || P < 1 || ld_new < N_new || ld_old < N_old) ABC_FAIL(ctx, ABCB200_EINVAL, "weights: bad shape");
    const int64_t ldn = pad32(N_new), ldo = pad32(N_old);
    size_t need = weights_ws_bytes(ctx, N_new, N_old, P) + align_up((size_t)ldn * P * 8, 256) + align_up((size_t)ldo * P * 8, 256) +
                  3 * align_up((size_t)N_new * 8, 256) + align_up((size_t)N_old * 8, 256) + align_up((size_t)P * 8, 256) + 4096;
    ABC_TRY(ws_reserve(ctx, need));
    double* d_new = ws_new<double>(ctx, (size_t)ldn * P);
    double* d_old = ws_new<double>(ctx, (size_t)ldo * P);
    double* d_numer = numer ? ws_new<double>(ctx, N_new) : nullptr;
    double* d_w = ws_new<double>(ctx, N_new);
    double* d_wold = ws_new<double>(ctx, N_old);
    double* d_dv = ws_new<double>(ctx, P);
    double* d_ss = ws_new<double>(ctx, 1);
    if (!d_new || !d_old || !d_w || !d_wold || !d_dv || !d_ss) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in weights");
    stage_begin(ctx, 8);
    ABC_TRY(h2d_matrix(ctx, d_new, ldn, theta_new, ld_new, N_new, P));
    ABC_TRY(h2d_matrix(ctx, d_old, ldo, theta_old, ld_old, N_old, P));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_wold, w_old, sizeof(double) * N_old, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_dv, dv_old, sizeof(double) * P, cudaMemcpyHostToDevice, ctx->stream));
    if (numer) CUDA_TRY(ctx, cudaMemcpyAsync(d_numer, numer, sizeof(double) * N_new, cudaMemcpyHostToDevice, ctx->stream));
    stage_end(ctx, 8);
    stage_begin(ctx, 7);
    ABC_TRY(weights_unnorm_dev(ctx, d_numer, d_new, ldn, N_new, d_old, ldo, N_old, d_wold, d_dv, P, algo, d_w, d_ss));
    ABC_TRY(launch_scale_weights(ctx, d_w, N_new, d_ss));
    stage_end(ctx, 7);
    stage_begin(ctx, 9);
    ABC_TRY(d2h(ctx, w_out, d_w, sizeof(double) * (size_t)N_new));
    stage_end(ctx, 9);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ABCB200_OK;
}

// ---- free functions -------------------------------------------------------------------------------------------
extern "C" int abcb200_colwise_moments(abcb200_ctx* ctx, const double* X, int64_t ld, int64_t N, int K, double* mean_out, double* sd_out) {
    ABC_TRY(check_ctx(ctx));
    if (!X || N < 1 || K < 1 || ld < N) ABC_FAIL(ctx, ABCB200_EINVAL, "colwise_moments: bad argument");
    const int64_t ldd = pad32(N);
    ABC_TRY(ws_reserve(ctx, moments_ws_bytes(N, K) + align_up((size_t)ldd * K * 8, 256) + 2 * align_up((size_t)K * 8, 256) + 2048));
    double* d = ws_new<double>(ctx, (size_t)ldd * K);
    double* d_mean = ws_new<double>(ctx, K);
    double* d_sd = ws_new<double>(ctx, K);
    double* stats = (double*)ws_alloc(ctx, moments_ws_bytes(N, K));
    if (!d || !d_mean || !d_sd || !stats) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted");
    ABC_TRY(h2d_matrix(ctx, d, ldd, X, ld, N, K));
    int nchunk = 0;
    ABC_TRY(launch_col_stats(ctx, d, ldd, N, K, stats, &nchunk));
    ABC_TRY(launch_col_finalize(ctx, stats, nchunk, N, K, d_mean, d_sd, 1.0, nullptr));
    if (mean_out) ABC_TRY(d2h(ctx, mean_out, d_mean, sizeof(double) * K));
    if (sd_out) ABC_TRY(d2h(ctx, sd_out, d_sd, sizeof(double) * K));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ABCB200_OK;
}

extern "C" int abcb200_colwise_z_scores(abcb200_ctx* ctx, const double* X, int64_t ld, int64_t N, int K, const double* mean,
                                        const double* sd, double* Z_out, int64_t ld_out) {
    ABC_TRY(check_ctx(ctx));
    if (!X || !Z_out || N < 1 || K < 1 || ld < N || ld_out < N || ((mean == nullptr) != (sd == nullptr))) ABC_FAIL(ctx, ABCB200_EINVAL, "colwise_z_scores: bad argument");
    const int64_t ldd = pad32(N);
    ABC_TRY(ws_reserve(ctx, moments_ws_bytes(N, K) + 2 * align_up((size_t)ldd * K * 8, 256) + 2 * align_up((size_t)K * 8, 256) + 2048));
    double* d = ws_new<double>(ctx, (size_t)ldd * K);
    double* dz = ws_new<double>(ctx, (size_t)ldd * K);
    double* d_mean = ws_new<double>(ctx, K);
    double* d_sd = ws_new<double>(ctx, K);
    double* stats = (double*)ws_alloc(ctx, moments_ws_bytes(N, K));
    if (!d || !dz || !d_mean || !d_sd || !stats) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted");
    ABC_TRY(h2d_matrix(ctx, d, ldd, X, ld, N, K));
    if (mean) {
        CUDA_TRY(ctx, cudaMemcpyAsync(d_mean, mean, sizeof(double) * K, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(d_sd, sd, sizeof(double) * K, cudaMemcpyHostToDevice, ctx->stream));
        ABC_TRY(launch_zscore(ctx, d, ldd, N, K, nullptr, 0, d_mean, d_sd, dz, ldd, nullptr, nullptr, nullptr, nullptr));
    } else {
        int nchunk = 0;
        ABC_TRY(launch_col_stats(ctx, d, ldd, N, K, stats, &nchunk));
        ABC_TRY(launch_zscore(ctx, d, ldd, N, K, stats, nchunk, nullptr, nullptr, dz, ldd, nullptr, nullptr, nullptr, nullptr));
    }
    CUDA_TRY(ctx, cudaMemcpy2DAsync(Z_out, (size_t)ld_out * 8, dz, (size_t)ldd * 8, (size_t)N * 8, (size_t)K, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ABCB200_OK;
}

extern "C" int abcb200_gram(abcb200_ctx* ctx, const double* X, int64_t ldx, const double* Y, int64_t ldy, int64_t N, int K, int M,
                            double* XX_out, double* XY_out) {
    ABC_TRY(check_ctx(ctx));
    if (!X || !Y || !XX_out || !XY_out || N < 1 || K < 1 || M < 1 || ldx < N || ldy < N) ABC_FAIL(ctx, ABCB200_EINVAL, "gram: bad argument");
    const int64_t ldd = pad32(N);
    ABC_TRY(ws_reserve(ctx, align_up((size_t)ldd * K * 8, 256) + align_up((size_t)ldd * M * 8, 256) + align_up((size_t)K * K * 8, 256) +
                                align_up((size_t)K * M * 8, 256) + gram_ws_bytes(ctx, N, K, M) + 4096));
    double* dX = ws_new<double>(ctx, (size_t)ldd * K);
    double* dY = ws_new<double>(ctx, (size_t)ldd * M);
    double* dXX = ws_new<double>(ctx, (size_t)K * K);
    double* dXY = ws_new<double>(ctx, (size_t)K * M);
    if (!dX || !dY || !dXX || !dXY) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted");
    ABC_TRY(h2d_matrix(ctx, dX, ldd, X, ldx, N, K));
    ABC_TRY(h2d_matrix(ctx, dY, ldd, Y, ldy, N, M));
    ABC_TRY(launch_gram(ctx, dX, ldd, K, dY, ldd, M, N, dXX, dXY));
    ABC_TRY(d2h(ctx, XX_out, dXX, sizeof(double) * (size_t)K * K));
    ABC_TRY(d2h(ctx, XY_out, dXY, sizeof(double) * (size_t)K * M));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ABCB200_OK;
}

extern "C" int abcb200_euclidean(abcb200_ctx* ctx, const double* S, int64_t ld, int64_t N, int K, const double* ref, double* out) {
    ABC_TRY(check_ctx(ctx));
    if (!S || !ref || !out || N < 1 || K < 1 || ld < N) ABC_FAIL(ctx, ABCB200_EINVAL, "euclidean: bad argument");
    const int64_t ldd = pad32(N);
    ABC_TRY(ws_reserve(ctx, align_up((size_t)ldd * K * 8, 256) + align_up((size_t)K * 8, 256) + align_up((size_t)N * 8, 256) + 2048));
    double* d = ws_new<double>(ctx, (size_t)ldd * K);
    double* d_ref = ws_new<double>(ctx, K);
    double* d_out = ws_new<double>(ctx, N);
    if (!d || !d_ref || !d_out) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted");
    ABC_TRY(h2d_matrix(ctx, d, ldd, S, ld, N, K));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_ref, ref, sizeof(double) * K, cudaMemcpyHostToDevice, ctx->stream));
    ABC_TRY(launch_euclidean(ctx, d, ldd, N, K, d_ref, d_out));
    ABC_TRY(d2h(ctx, out, d_out, sizeof(double) * (size_t)N));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ABCB200_OK;
}

extern "C" int abcb200_ordered_top(abcb200_ctx* ctx, const double* v, int64_t n, int64_t top_n, uint64_t* order_out) {
    ABC_TRY(check_ctx(ctx));
    if (!v || !order_out || n < 0) ABC_FAIL(ctx, ABCB200_EINVAL, "ordered: bad argument");
    if (n == 0) return ABCB200_OK;
    if (top_n <= 0 || top_n > n) top_n = n;
    ABC_TRY(ws_reserve(ctx, order_ws_bytes(n) + 2 * align_up((size_t)n * 8, 256) + 1024));
    double* d = ws_new<double>(ctx, n);
    uint64_t* d_ord = ws_new<uint64_t>(ctx, top_n);
    if (!d || !d_ord) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted");
    CUDA_TRY(ctx, cudaMemcpyAsync(d, v, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    ABC_TRY(order_dev(ctx, d, n, top_n, d_ord));
    ABC_TRY(d2h(ctx, order_out, d_ord, sizeof(uint64_t) * (size_t)top_n));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->tie_order == 1 && tie_order_stdsort(v, n, top_n, order_out)) ctx->stat_tie_resorts++;
    return ABCB200_OK;
}
extern "C" int abcb200_ordered(abcb200_ctx* ctx, const double* v, int64_t n, uint64_t* order_out) {
    return abcb200_ordered_top(ctx, v, n, n, order_out);
}

extern "C" int abcb200_wilcoxon(abcb200_ctx* ctx, const double* err1, const double* err2, int64_t n, double* p_out) {
    ABC_TRY(check_ctx(ctx));
    if (!err1 || !err2 || !p_out || n < 1) ABC_FAIL(ctx, ABCB200_EINVAL, "wilcoxon: bad argument");
    ABC_TRY(ws_reserve(ctx, wilcoxon_ws_bytes(n) + 2 * align_up((size_t)n * 8, 256) + 1024));
    double* d1 = ws_new<double>(ctx, n);
    double* d2 = ws_new<double>(ctx, n);
    if (!d1 || !d2) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted");
    CUDA_TRY(ctx, cudaMemcpyAsync(d1, err1, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(d2, err2, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    return wilcoxon_dev(ctx, d1, d2, n, p_out);
}

// ---- PLS::Model ------------------------------------------------------------------------------------------------
extern "C" int abcb200_pls_fit(abcb200_ctx* ctx, const double* X, int64_t ldx, const double* Y, int64_t ldy, int64_t N, int K, int M,
                               int method, int max_components, abcb200_pls** out) {
    ABC_TRY(check_ctx(ctx));
    if (!out) return ABCB200_EINVAL;
    *out = nullptr;
    if (!X || !Y || N < 1 || K < 1 || M < 1 || ldx < N || ldy < N) ABC_FAIL(ctx, ABCB200_EINVAL, "pls_fit: bad argument");
    if (max_components < 1 || max_components > K) ABC_FAIL(ctx, ABCB200_EINVAL, "pls_fit: max_components %d outside [1, K=%d] (pls.cpp:345)", max_components, K);
    if (method < ABCB200_KERNEL_TYPE1 || method > ABCB200_KERNEL_TYPE1_STREAM) ABC_FAIL(ctx, ABCB200_EINVAL, "pls_fit: unknown method %d", method);
    const int A = max_components;
    const int64_t ldd = pad32(N);
    abcb200_pls* m = new (std::nothrow) abcb200_pls();
    if (!m) return ABCB200_ENOMEM;
    m->ctx = ctx;
    const bool keepT = method != ABCB200_KERNEL_TYPE2;
    const size_t nd = (size_t)3 * K * A + (size_t)M * A + (keepT ? (size_t)ldd * A : 0);
    if (cudaMalloc(&m->storage, nd * sizeof(double)) != cudaSuccess) { cudaGetLastError(); delete m; ABC_FAIL(ctx, ABCB200_ENOMEM, "pls_fit: model storage"); }
    PlsFactors& f = m->f;
    f.K = K; f.M = M; f.A = A; f.method = method; f.n = N;
    f.W = m->storage; f.P = f.W + (size_t)K * A; f.R = f.P + (size_t)K * A; f.Q = f.R + (size_t)K * A;
    f.T = keepT ? f.Q + (size_t)M * A : nullptr; f.ldt = ldd;
    int rc = ws_reserve(ctx, pls_fit_ws_bytes(ctx, N, K, M, method) + align_up((size_t)ldd * K * 8, 256) + align_up((size_t)ldd * M * 8, 256) + 4096);
    double *dX = nullptr, *dY = nullptr;
    if (rc == ABCB200_OK) {
        dX = ws_new<double>(ctx, (size_t)ldd * K);
        dY = ws_new<double>(ctx, (size_t)ldd * M);
        if (!dX || !dY) rc = ABCB200_ENOMEM;
    }
    if (rc == ABCB200_OK) rc = h2d_matrix(ctx, dX, ldd, X, ldx, N, K);
    if (rc == ABCB200_OK) rc = h2d_matrix(ctx, dY, ldd, Y, ldy, N, M);
    if (rc == ABCB200_OK) { stage_begin(ctx, 1); rc = pls_fit_dev(ctx, dX, ldd, dY, ldd, f); stage_end(ctx, 1); }
    if (rc == ABCB200_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = ABCB200_ECUDA;
    if (rc != ABCB200_OK) { cudaFree(m->storage); delete m; return rc; }
    *out = m;
    return ABCB200_OK;
}

extern "C" int abcb200_pls_free(abcb200_pls* m) {
    if (!m) return ABCB200_OK;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    cudaFree(m->storage);
    delete m;
    return ABCB200_OK;
}

extern "C" int abcb200_pls_get(abcb200_pls* m, char which, double* out) {
    if (!m || !out) return ABCB200_EINVAL;
    abcb200_ctx* ctx = m->ctx;
    ABC_TRY(check_ctx(ctx));
    const PlsFactors& f = m->f;
    switch (which) {
        case 'W': ABC_TRY(d2h(ctx, out, f.W, sizeof(double) * f.K * f.A)); break;
        case 'P': ABC_TRY(d2h(ctx, out, f.P, sizeof(double) * f.K * f.A)); break;
        case 'R': ABC_TRY(d2h(ctx, out, f.R, sizeof(double) * f.K * f.A)); break;
        case 'Q': ABC_TRY(d2h(ctx, out, f.Q, sizeof(double) * f.M * f.A)); break;
        case 'T':
            if (!f.T) ABC_FAIL(ctx, ABCB200_EINVAL, "pls_get: T exists only for KERNEL_TYPE1 (pls.cpp:394)");
            CUDA_TRY(ctx, cudaMemcpy2DAsync(out, (size_t)f.n * 8, f.T, (size_t)f.ldt * 8, (size_t)f.n * 8, (size_t)f.A, cudaMemcpyDeviceToHost, ctx->stream));
            break;
        default: ABC_FAIL(ctx, ABCB200_EINVAL, "pls_get: unknown matrix '%c'", which);
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ABCB200_OK;
}

// shared driver: out = op(Xnew [, Ynew]) with comp components. mode 0 scores, 1 fitted, 2 residuals, 3 SSE
static int pls_apply(abcb200_pls* m, const double* Xn, int64_t ldx, const double* Yn, int64_t ldy, int64_t n, int comp, int mode, double* out) {
    if (!m) return ABCB200_EINVAL;
    abcb200_ctx* ctx = m->ctx;
    ABC_TRY(check_ctx(ctx));
    const PlsFactors& f = m->f;
    if (!Xn || !out || n < 1 || ldx < n || comp < 0 || comp > f.A) ABC_FAIL(ctx, ABCB200_EINVAL, "pls: bad argument (comp=%d, A=%d; pls.cpp:440)", comp, f.A);
    if (mode >= 2 && (!Yn || ldy < n)) ABC_FAIL(ctx, ABCB200_EINVAL, "pls: Y required");
    const int64_t ldd = pad32(n);
    const int ncols = mode == 0 ? comp : f.M;
    size_t need = align_up((size_t)ldd * f.K * 8, 256) + 2 * align_up((size_t)ldd * (size_t)(ncols > 0 ? ncols : 1) * 8, 256) + align_up((size_t)ldd * f.M * 8, 256) +
                  align_up((size_t)f.K * f.M * 8, 256) + align_up((size_t)f.M * 8, 256) + 4096;
    ABC_TRY(ws_reserve(ctx, need));
    double* dX = ws_new<double>(ctx, (size_t)ldd * f.K);
    double* dO = ws_new<double>(ctx, (size_t)ldd * (ncols > 0 ? ncols : 1));
    if (!dX || !dO) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted");
    ABC_TRY(h2d_matrix(ctx, dX, ldd, Xn, ldx, n, f.K));
    if (mode == 0) {
        if (comp == 0) return ABCB200_OK;
        ABC_TRY(launch_xb(ctx, dX, ldd, n, f.K, f.R, f.K, comp, dO, ldd));
        CUDA_TRY(ctx, cudaMemcpy2DAsync(out, (size_t)n * 8, dO, (size_t)ldd * 8, (size_t)n * 8, (size_t)comp, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        double* dB = ws_new<double>(ctx, (size_t)f.K * f.M);
        if (!dB) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted");
        ABC_TRY(launch_coefficients(ctx, f.R, f.Q, f.K, f.M, comp, dB));
        ABC_TRY(launch_xb(ctx, dX, ldd, n, f.K, dB, f.K, f.M, dO, ldd));
        if (mode >= 2) {
            double* dY = ws_new<double>(ctx, (size_t)ldd * f.M);
            double* dE = ws_new<double>(ctx, (size_t)ldd * f.M);
            double* dS = ws_new<double>(ctx, f.M);
            if (!dY || !dE || !dS) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted");
            ABC_TRY(h2d_matrix(ctx, dY, ldd, Yn, ldy, n, f.M));
            const int gx = (int)std::max((int64_t)1, std::min((n + 255) / 256, (int64_t)1024));
            LAUNCH(ctx, residual_kernel, dim3(gx, f.M), 256, 0, dY, ldd, dO, ldd, n, f.M, dE, ldd);
            if (mode == 3) {
                LAUNCH(ctx, colsumsq_kernel, f.M, 256, 0, dE, ldd, n, dS);
                ABC_TRY(d2h(ctx, out, dS, sizeof(double) * f.M));
            } else {
                CUDA_TRY(ctx, cudaMemcpy2DAsync(out, (size_t)n * 8, dE, (size_t)ldd * 8, (size_t)n * 8, (size_t)f.M, cudaMemcpyDeviceToHost, ctx->stream));
            }
        } else {
            CUDA_TRY(ctx, cudaMemcpy2DAsync(out, (size_t)n * 8, dO, (size_t)ldd * 8, (size_t)n * 8, (size_t)f.M, cudaMemcpyDeviceToHost, ctx->stream));
        }
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ABCB200_OK;
}

extern "C" int abcb200_pls_scores(abcb200_pls* m, const double* Xnew, int64_t ld, int64_t n, int comp, double* out) {
    return pls_apply(m, Xnew, ld, nullptr, 0, n, comp, 0, out);
}
extern "C" int abcb200_pls_fitted_values(abcb200_pls* m, const double* Xnew, int64_t ld, int64_t n, int comp, double* out) {
    return pls_apply(m, Xnew, ld, nullptr, 0, n, comp, 1, out);
}
extern "C" int abcb200_pls_residuals(abcb200_pls* m, const double* Xnew, int64_t ldx, const double* Ynew, int64_t ldy, int64_t n, int comp, double* out) {
    return pls_apply(m, Xnew, ldx, Ynew, ldy, n, comp, 2, out);
}
extern "C" int abcb200_pls_sse(abcb200_pls* m, const double* Xnew, int64_t ldx, const double* Ynew, int64_t ldy, int64_t n, int comp, double* out) {
    return pls_apply(m, Xnew, ldx, Ynew, ldy, n, comp, 3, out);
}

/* Model::explained_variance, pls.cpp:461-467: 1 - SSE / SST(Y_new) */
extern "C" int abcb200_pls_explained_variance(abcb200_pls* m, const double* Xnew, int64_t ldx, const double* Ynew, int64_t ldy, int64_t n, int comp, double* out) {
    ABC_TRY(pls_apply(m, Xnew, ldx, Ynew, ldy, n, comp, 3, out));
    const int M = m->f.M;
    std::vector<double> mean(M), sd(M);
    ABC_TRY(abcb200_colwise_moments(m->ctx, Ynew, ldy, n, M, mean.data(), sd.data()));
    for (int y = 0; y < M; y++) out[y] = 1.0 - out[y] / (sd[y] * sd[y] * (double)(n - 1));     // SST = sd^2 (n - 1), pls.cpp:69-87
    return ABCB200_OK;
}

extern "C" int abcb200_pls_coefficients(abcb200_pls* m, int comp, double* out) {
    if (!m || !out) return ABCB200_EINVAL;
    abcb200_ctx* ctx = m->ctx;
    ABC_TRY(check_ctx(ctx));
    const PlsFactors& f = m->f;
    if (comp < 0 || comp > f.A) ABC_FAIL(ctx, ABCB200_EINVAL, "pls_coefficients: comp=%d > A=%d (pls.cpp:445)", comp, f.A);
    ABC_TRY(ws_reserve(ctx, align_up((size_t)f.K * f.M * 8, 256) + 1024));
    double* dB = ws_new<double>(ctx, (size_t)f.K * f.M);
    ABC_TRY(launch_coefficients(ctx, f.R, f.Q, f.K, f.M, comp, dB));
    ABC_TRY(d2h(ctx, out, dB, sizeof(double) * f.K * f.M));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ABCB200_OK;
}

extern "C" int abcb200_pls_cv_new_data(abcb200_pls* m, const double* Xnew, int64_t ldx, const double* Ynew, int64_t ldy, int64_t n,
                                       int out_type, double alpha, double* press_out, int32_t* n_comp_out) {
    if (!m) return ABCB200_EINVAL;
    abcb200_ctx* ctx = m->ctx;
    ABC_TRY(check_ctx(ctx));
    const PlsFactors& f = m->f;
    if (!Xnew || !Ynew || n < 1 || ldx < n || ldy < n) ABC_FAIL(ctx, ABCB200_EINVAL, "pls_cv_new_data: bad argument");
    if (f.M > 128) ABC_FAIL(ctx, ABCB200_EINVAL, "pls_cv_new_data: M > 128");
    const int64_t ldd = pad32(n);
    ABC_TRY(ws_reserve(ctx, holdout_ws_bytes(ctx, n, f.K, f.M, f.A) + align_up((size_t)ldd * f.K * 8, 256) + align_up((size_t)ldd * f.M * 8, 256) +
                                align_up((size_t)f.M * f.A * 8, 256) + 4096));
    double* dX = ws_new<double>(ctx, (size_t)ldd * f.K);
    double* dY = ws_new<double>(ctx, (size_t)ldd * f.M);
    double* dPress = ws_new<double>(ctx, (size_t)f.M * f.A);
    if (!dX || !dY || !dPress) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted");
    ABC_TRY(h2d_matrix(ctx, dX, ldd, Xnew, ldx, n, f.K));
    ABC_TRY(h2d_matrix(ctx, dY, ldd, Ynew, ldy, n, f.M));
    ABC_TRY(holdout_select_dev(ctx, dX, ldd, dY, ldd, n, f, alpha, dPress, n_comp_out));
    if (press_out) {
        ABC_TRY(d2h(ctx, press_out, dPress, sizeof(double) * f.M * f.A));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (out_type == ABCB200_MSE) for (size_t i = 0; i < (size_t)f.M * f.A; i++) press_out[i] /= (double)n;   // pls.cpp:257
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ABCB200_OK;
}

// ---- validation from a residual cube, Model::cv_LOO, Model::cv_LSO -------------------------------------------------------
static int cube_to_host(abcb200_ctx* ctx, double* errors_out, const double* cube, size_t count) {
    if (!errors_out || count == 0) return ABCB200_OK;
    ABC_TRY(d2h(ctx, errors_out, cube, sizeof(double) * count));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ABCB200_OK;
}

extern "C" int abcb200_residual_select(abcb200_ctx* ctx, const double* errors, int64_t n, int M, int A, int out_type, double alpha,
                                       double* press_out, int32_t* n_comp_out) {
    ABC_TRY(check_ctx(ctx));
    if (!errors || n < 1 || M < 1 || A < 1) ABC_FAIL(ctx, ABCB200_EINVAL, "residual_select: bad argument");
    const size_t count = (size_t)n * M * A;
    ABC_TRY(ws_reserve(ctx, align_up(count * 8, 256) + cube_select_ws_bytes(n, M, A) + 4096));
    double* cube = ws_new<double>(ctx, count);
    if (!cube) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted");
    CUDA_TRY(ctx, cudaMemcpyAsync(cube, errors, count * 8, cudaMemcpyHostToDevice, ctx->stream));
    stage_begin(ctx, 3);
    const int rc = cube_select_dev(ctx, cube, n, M, A, out_type, alpha, press_out, n_comp_out);
    stage_end(ctx, 3);
    return rc;
}

extern "C" int abcb200_pls_cv_loo(abcb200_ctx* ctx, const double* X, int64_t ldx, const double* Y, int64_t ldy, int64_t N, int K, int M,
                                  int A, int out_type, double alpha, double* errors_out, double* press_out, int32_t* n_comp_out) {
    ABC_TRY(check_ctx(ctx));
    if (!X || !Y || N < 2 || K < 1 || M < 1 || ldx < N || ldy < N) ABC_FAIL(ctx, ABCB200_EINVAL, "pls_cv_loo: bad argument");
    if (A < 1 || A > K) ABC_FAIL(ctx, ABCB200_EINVAL, "pls_cv_loo: A=%d outside [1, K=%d]", A, K);
    if (M > 128) ABC_FAIL(ctx, ABCB200_EINVAL, "pls_cv_loo: M > 128");
    const int64_t ldd = pad32(N);
    const size_t count = (size_t)N * M * A;
    const bool on_chip = pls_loo_fits(ctx, K, M);
    size_t need = align_up((size_t)ldd * K * 8, 256) + align_up((size_t)ldd * M * 8, 256) + 2 * align_up((size_t)K * K * 8, 256) +
                  2 * align_up((size_t)K * M * 8, 256) + gram_ws_bytes(ctx, N, K, M) + align_up(count * 8, 256) + cube_select_ws_bytes(N, M, A) + 8192;
    if (!on_chip) need += 3 * align_up((size_t)K * A * 8, 256) + 2 * align_up((size_t)M * A * 8, 256) + align_up((size_t)A * 8, 256) + pls_components_ws_bytes(K, M, A);
    ABC_TRY(ws_reserve(ctx, need));
    double* dX = ws_new<double>(ctx, (size_t)ldd * K);
    double* dY = ws_new<double>(ctx, (size_t)ldd * M);
    double* XX = ws_new<double>(ctx, (size_t)K * K);
    double* XY = ws_new<double>(ctx, (size_t)K * M);
    double* cube = ws_new<double>(ctx, count);
    if (!dX || !dY || !XX || !XY || !cube) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted");
    ABC_TRY(h2d_matrix(ctx, dX, ldd, X, ldx, N, K));
    ABC_TRY(h2d_matrix(ctx, dY, ldd, Y, ldy, N, M));
    stage_begin(ctx, 1);
    ABC_TRY(launch_gram(ctx, dX, ldd, K, dY, ldd, M, N, XX, XY));
    if (on_chip) {
        ABC_TRY(pls_loo_dev(ctx, dX, ldd, dY, ldd, N, K, M, A, XX, XY, cube));
    } else {
        // wide predictor blocks: the down-dated Gram matrices go through the L2-streamed component loop, one held-out row at a time
        double* XXi = ws_new<double>(ctx, (size_t)K * K);
        double* XYi = ws_new<double>(ctx, (size_t)K * M);
        double* fac = ws_new<double>(ctx, (size_t)3 * K * A + (size_t)M * A);
        double* trow = ws_new<double>(ctx, A);
        if (!XXi || !XYi || !fac || !trow) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted");
        PlsFactors f;
        f.K = K; f.M = M; f.A = A; f.method = ABCB200_KERNEL_TYPE2; f.n = N - 1; f.T = nullptr; f.ldt = 0;
        f.W = fac; f.P = fac + (size_t)K * A; f.R = f.P + (size_t)K * A; f.Q = f.R + (size_t)K * A;
        const size_t mark = ctx->ws_off;
        for (int64_t i = 0; i < N; i++) {
            ctx->ws_off = mark;
            ABC_TRY(launch_gram_downdate(ctx, XX, XY, dX, ldd, dY, ldd, i, K, M, XXi, XYi));
            ABC_TRY(pls_components_dev(ctx, XXi, XYi, f));
            ABC_TRY(launch_xb(ctx, dX + i, ldd, 1, K, f.R, K, A, trow, 1));
            ABC_TRY(launch_cube_from_scores(ctx, trow, 1, dY + i, ldd, f.Q, 1, M, A, cube, N, i, false));
        }
        ctx->ws_off = mark;
    }
    stage_end(ctx, 1);
    ABC_TRY(cube_to_host(ctx, errors_out, cube, count));
    stage_begin(ctx, 3);
    const int rc = (press_out || n_comp_out) ? cube_select_dev(ctx, cube, N, M, A, out_type, alpha, press_out, n_comp_out) : ABCB200_OK;
    stage_end(ctx, 3);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return rc;
}

extern "C" int abcb200_pls_cv_lso(abcb200_ctx* ctx, const double* X, int64_t ldx, const double* Y, int64_t ldy, int64_t N, int K, int M,
                                  int A, int method, const uint64_t* shuffles, int64_t test_size, int64_t num_trials, int out_type,
                                  double alpha, double* errors_out, double* press_out, int32_t* n_comp_out) {
    ABC_TRY(check_ctx(ctx));
    if (!X || !Y || !shuffles || N < 2 || K < 1 || M < 1 || ldx < N || ldy < N || num_trials < 1) ABC_FAIL(ctx, ABCB200_EINVAL, "pls_cv_lso: bad argument");
    if (A < 1 || A > K) ABC_FAIL(ctx, ABCB200_EINVAL, "pls_cv_lso: A=%d outside [1, K=%d]", A, K);
    if (test_size < 1 || test_size >= N) ABC_FAIL(ctx, ABCB200_EINVAL, "pls_cv_lso: test_size=%lld must leave both parts non-empty (pls.cpp:518)", (long long)test_size);
    if (method < ABCB200_KERNEL_TYPE1 || method > ABCB200_KERNEL_TYPE1_STREAM) ABC_FAIL(ctx, ABCB200_EINVAL, "pls_cv_lso: unknown method %d", method);
    for (int64_t i = 0; i < num_trials * N; i++)
        if (shuffles[i] >= (uint64_t)N) ABC_FAIL(ctx, ABCB200_EINVAL, "pls_cv_lso: shuffle entry %lld out of range", (long long)i);
    const int64_t n_tr = N - test_size, rows = num_trials * test_size;
    const int64_t ldd = pad32(N), ldt = pad32(n_tr), ldp = pad32(test_size);
    const size_t count = (size_t)rows * M * A;
    ABC_TRY(ws_reserve(ctx, align_up((size_t)ldd * (K + M) * 8, 512) + align_up((size_t)ldt * (K + M) * 8, 512) + align_up((size_t)ldp * (K + M + A) * 8, 768) +
                                align_up((size_t)N * 8, 256) + align_up(((size_t)3 * K * A + (size_t)M * A) * 8, 256) + pls_fit_ws_bytes(ctx, n_tr, K, M, method) +
                                align_up(count * 8, 256) + cube_select_ws_bytes(rows, M, A) + 16384));
    double* dX = ws_new<double>(ctx, (size_t)ldd * K);
    double* dY = ws_new<double>(ctx, (size_t)ldd * M);
    double* dXv = ws_new<double>(ctx, (size_t)ldt * K);
    double* dYv = ws_new<double>(ctx, (size_t)ldt * M);
    double* dXp = ws_new<double>(ctx, (size_t)ldp * K);
    double* dYp = ws_new<double>(ctx, (size_t)ldp * M);
    double* dTp = ws_new<double>(ctx, (size_t)ldp * A);
    uint64_t* didx = ws_new<uint64_t>(ctx, N);
    double* fac = ws_new<double>(ctx, (size_t)3 * K * A + (size_t)M * A);
    double* cube = ws_new<double>(ctx, count);
    if (!dX || !dY || !dXv || !dYv || !dXp || !dYp || !dTp || !didx || !fac || !cube) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted");
    ABC_TRY(h2d_matrix(ctx, dX, ldd, X, ldx, N, K));
    ABC_TRY(h2d_matrix(ctx, dY, ldd, Y, ldy, N, M));
    PlsFactors f;
    f.K = K; f.M = M; f.A = A; f.method = method; f.n = n_tr; f.T = nullptr; f.ldt = 0;
    f.W = fac; f.P = fac + (size_t)K * A; f.R = f.P + (size_t)K * A; f.Q = f.R + (size_t)K * A;
    const size_t mark = ctx->ws_off;
    stage_begin(ctx, 1);
    for (int64_t rep = 0; rep < num_trials; rep++) {
        ctx->ws_off = mark;
        // sample = full[0, n_tr), complement = full[n_tr, N) (rand_nchoosek, pls.cpp:218-227)
        CUDA_TRY(ctx, cudaMemcpyAsync(didx, shuffles + rep * N, sizeof(uint64_t) * (size_t)N, cudaMemcpyHostToDevice, ctx->stream));
        ABC_TRY(launch_gather_rows(ctx, dX, ldd, didx, n_tr, K, dXv, ldt));
        ABC_TRY(launch_gather_rows(ctx, dY, ldd, didx, n_tr, M, dYv, ldt));
        ABC_TRY(launch_gather_rows(ctx, dX, ldd, didx + n_tr, test_size, K, dXp, ldp));
        ABC_TRY(launch_gather_rows(ctx, dY, ldd, didx + n_tr, test_size, M, dYp, ldp));
        ABC_TRY(pls_fit_dev(ctx, dXv, ldt, dYv, ldt, f));                                            // pls.cpp:539
        ABC_TRY(launch_xb(ctx, dXp, ldp, test_size, K, f.R, K, A, dTp, ldp));
        ABC_TRY(launch_cube_from_scores(ctx, dTp, ldp, dYp, ldp, f.Q, test_size, M, A, cube, rows, rep * test_size, false));   // :540-545 (+= into zeros)
    }
    ctx->ws_off = mark;
    stage_end(ctx, 1);
    ABC_TRY(cube_to_host(ctx, errors_out, cube, count));
    stage_begin(ctx, 3);
    const int rc = (press_out || n_comp_out) ? cube_select_dev(ctx, cube, rows, M, A, out_type, alpha, press_out, n_comp_out) : ABCB200_OK;
    stage_end(ctx, 3);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return rc;
}
