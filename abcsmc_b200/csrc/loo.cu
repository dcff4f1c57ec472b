// loo.cu — validation and component selection from a MATERIALISED residual cube (SURVEY.md §8 rows a7, a8, a10'').
//
// Reference: PLS::validation (lib/PLS/src/pls.cpp:235-261) and PLS::optimal_num_components (:265-289) take a
// PLS::Residual = one (rows x A) error matrix per response. The hold-out path of AbcSmc never materialises that cube
// (holdout.cu); Model::cv_LOO (:469-491) and Model::cv_LSO (:512-549) produce it row block by row block, so here it lives
// in HBM as cube[(y * A + c) * n + i]: every error column the Wilcoxon test reads is one contiguous stream.
#include <vector>

#include "kernels.cuh"

namespace {

// press[y * A + c] = sum_i cube[(y * A + c) * n + i]^2 (pls.cpp:247-256). One CTA per column, fixed summation order.
__global__ void __launch_bounds__(256) cube_press_kernel(const double* __restrict__ cube, long long n, double* __restrict__ press) {
    __shared__ double red[32];
    const double* col = cube + (size_t)blockIdx.x * (size_t)n;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    long long i = threadIdx.x;
    for (; i + 768 < n; i += 1024) {
        const double v0 = col[i], v1 = col[i + 256], v2 = col[i + 512], v3 = col[i + 768];
        a0 = fma(v0, v0, a0); a1 = fma(v1, v1, a1); a2 = fma(v2, v2, a2); a3 = fma(v3, v3, a3);
    }
    for (; i < n; i += 256) { const double v = col[i]; a0 = fma(v, v, a0); }
    const double s = block_sum((a0 + a1) + (a2 + a3), red);
    if (threadIdx.x == 0) press[blockIdx.x] = s;
}

// errors of new rows from their scores: cube[(y * A + c) * n + i] = Y[i, y] - sum_{a <= c} T[i, a] Q[y, a]
// (Model::residuals for c + 1 components, pls.cpp:449-455, as the prefix recurrence e_c = e_{c-1} - t_c q_c).
// add != 0 accumulates into the cube (cv_LSO's "+=", pls.cpp:543).
__global__ void __launch_bounds__(256) cube_from_scores_kernel(const double* __restrict__ T, long long ldt, const double* __restrict__ Y, long long ldy,
                                                               const double* __restrict__ Q, long long n, int M, int A, double* __restrict__ cube,
                                                               long long cube_n, long long row0, int add) {
    const int y = blockIdx.y;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += 256LL * gridDim.x) {
        double e = Y[(size_t)y * ldy + i];
        for (int c = 0; c < A; c++) {
            e = fma(-T[(size_t)c * ldt + i], Q[(size_t)c * M + y], e);
            double* dst = cube + ((size_t)y * A + c) * (size_t)cube_n + (size_t)(row0 + i);
            *dst = add ? *dst + e : e;
        }
    }
}

// leave-one-out Gram matrices of row `row`: XXo = XX - x x^T, XYo = XY - x y^T (rank-one down-dates)
__global__ void __launch_bounds__(256) gram_downdate_kernel(const double* __restrict__ XX, const double* __restrict__ XY, const double* __restrict__ X,
                                                            long long ldx, const double* __restrict__ Y, long long ldy, long long row, int K, int M,
                                                            double* __restrict__ XXo, double* __restrict__ XYo) {
    const long long tot = (long long)K * (K + M);
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < tot; i += 256LL * gridDim.x) {
        const int c = (int)(i / K), r = (int)(i - (long long)c * K);
        const double xr = X[(size_t)r * ldx + row];
        if (c < K) XXo[i] = fma(-xr, X[(size_t)c * ldx + row], XX[i]);
        else XYo[i - (long long)K * K] = fma(-xr, Y[(size_t)(c - K) * ldy + row], XY[i - (long long)K * K]);
    }
}

}  // namespace

int launch_gram_downdate(abcb200_ctx* ctx, const double* XX, const double* XY, const double* X, int64_t ldx, const double* Y, int64_t ldy,
                         int64_t row, int K, int M, double* XXo, double* XYo) {
    const long long tot = (long long)K * (K + M);
    const int grid = (int)std::max<long long>(1, std::min<long long>((tot + 255) / 256, 4 * ctx->sm_count));
    LAUNCH(ctx, gram_downdate_kernel, grid, 256, 0, XX, XY, X, (long long)ldx, Y, (long long)ldy, (long long)row, K, M, XXo, XYo);
    return ABCB200_OK;
}

size_t cube_select_ws_bytes(int64_t n, int M, int A) { return align_up((size_t)M * A * 8, 256) + wilcoxon_ws_bytes(n) + 1024; }

int launch_cube_from_scores(abcb200_ctx* ctx, const double* T, int64_t ldt, const double* Y, int64_t ldy, const double* Q, int64_t n, int M, int A,
                            double* cube, int64_t cube_n, int64_t row0, bool add) {
    if (n <= 0) return ABCB200_OK;
    const int gx = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 4 * ctx->sm_count));
    LAUNCH(ctx, cube_from_scores_kernel, dim3(gx, M), 256, 0, T, (long long)ldt, Y, (long long)ldy, Q, (long long)n, M, A, cube, (long long)cube_n,
           (long long)row0, add ? 1 : 0);
    return ABCB200_OK;
}

// PLS::validation + PLS::optimal_num_components on a device cube (M x A columns of n rows). press_host: M x A column-major
// [y + M * c] like the reference's Mat2D (nullable); ncomp_host: M counts (nullable).
int cube_select_dev(abcb200_ctx* ctx, const double* cube, int64_t n, int M, int A, int out_type, double alpha, double* press_host,
                    int32_t* ncomp_host) {
    double* dpress = ws_new<double>(ctx, (size_t)M * A);
    if (!dpress) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in cube_select");
    std::vector<double> press((size_t)M * A, 0.0);
    if (n > 0) {
        LAUNCH(ctx, cube_press_kernel, M * A, 256, 0, cube, (long long)n, dpress);
        CUDA_TRY(ctx, cudaMemcpyAsync(press.data(), dpress, sizeof(double) * M * A, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (press_host) {
        for (int y = 0; y < M; y++)
            for (int c = 0; c < A; c++) press_host[y + (size_t)M * c] = press[(size_t)y * A + c] / (out_type == ABCB200_MSE ? (double)n : 1.0);   // pls.cpp:257
    }
    if (!ncomp_host) return ABCB200_OK;
    const size_t ws_mark = ctx->ws_off;
    for (int y = 0; y < M; y++) {
        int ref = 0;                                                              // first minimum (Eigen minCoeff(&idx), pls.cpp:278)
        for (int c = 1; c < A; c++) if (press[(size_t)y * A + c] < press[(size_t)y * A + ref]) ref = c;
        const double* eref = cube + ((size_t)y * A + ref) * (size_t)n;
        for (int alt = 0; alt < ref && n > 0; alt++) {                            // smallest acceptable alternative (pls.cpp:281-286)
            double p = 0.0;
            ctx->ws_off = ws_mark;
            ABC_TRY(wilcoxon_dev(ctx, eref, cube + ((size_t)y * A + alt) * (size_t)n, n, &p));
            ctx->exact_tests++;
            if (p > alpha) { ref = alt; break; }
        }
        ncomp_host[y] = ref + 1;
    }
    ctx->ws_off = ws_mark;
    return ABCB200_OK;
}
