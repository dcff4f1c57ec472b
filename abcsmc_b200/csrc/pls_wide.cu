// pls_wide.cu — the component loop of kernel PLS for WIDE predictor sets (K > ~170: the deflated Gram matrix no longer
// fits one SM's shared memory), SURVEY.md §8 row a4 at the PLS-heavy shape (K = 500, M = 50).
//
// Reference: PLS::Model::plsr, lib/PLS/src/pls.cpp:400-435. Same formulation as pls_defl.cu — per component the dominant
// eigenvector q of XY^T XY (:406), w^ = XY q (:408), and with the DEFLATED Gram matrix H_a = X_a^T X_a (H_0 = X^T X):
//       p^ = H_a w^,  tt^ = w^^T p^,  q^ = XY_a^T w^,   H_{a+1} = H_a - p^ p^^T / tt^,   XY_{a+1} = XY_a - p^ q^^T / tt^ (:429)
// and W = w^/|w^|, P = p^ |w^|/tt^, Q = q^ |w^|/tt^ (:411, :427, :428); R follows from the reference's recurrence (:412-416,
// pls_u_kernel / pls_r_kernel of pls_defl.cu) — but H (K^2 doubles, 2 MB at K = 500) and XY live in L2 and the O(K^2) and
// O(K M^2) parts of a component run on the WHOLE GPU. pls_gram.cu streams those operands through one SM (measured 222k
// cycles = 113 us per component at K = 500: 62k for r = w - R P^T w, 78k for p = XX r, 33k for XY^T XY); here a component
// is three stream-ordered launches, no grid barrier, no atomics, every sum in a fixed order (deterministic):
//
//   wide_s0_kernel    one warp per entry (i <= j) of S0 = XY_a^T XY_a; the deflation of XY by the PREVIOUS component is applied
//                     on the fly (XY_{a-1} is only read, the warps of the diagonal entries write XY_a to the other buffer of
//                     a ping-pong pair); CTA 0 also emits W, P, Q of the previous component.
//   wide_eig_kernel   one CTA: projector onto the dominant eigenvector by trace-normalised repeated squaring on DMMA (the
//                     iteration of pls_defl.cu), then w^ = XY_a q and q^ = XY_a^T w^. This is the latency-bound part that remains.
//   wide_hw_kernel    one warp per row of H: applies the pending rank-one term of the previous component (symmetric product
//                     p_i p_j first, so H stays bitwise symmetric), stores the row, p^_i = H_a[i, :] w^.
//
// 1 / tt^ of the previous component is recomputed by every warp that needs it from w^ and p^ (K terms, same order, same
// bits everywhere) instead of being exchanged.
#include "kernels.cuh"

namespace {

constexpr int WE_T = 512;            // threads of the eigen CTA
constexpr int WE_W = WE_T / 32;

struct WideArgs {
    int K, M, A, comp;
    const double* XYold;             // K x M (ld K): XY_{a-1} (component 0: the input XY)
    double* XYnew;                   // XY_a
    double* H;                       // K x K, symmetric, updated in place
    const double* wprev;             // w^ of component a-1
    const double* pprev;             // p^ of component a-1
    const double* qprev;             // q^ of component a-1
    double* wcur;                    // w^ of component a
    double* pcur;                    // p^ of component a
    double* qcur;                    // q^ of component a
    double* S0;                      // M x M (ld M), full symmetric
    double *W, *P, *Q;               // outputs, K x A / K x A / M x A
};

__device__ __forceinline__ double pow2_inv_w(double x) {     // 2^-exponent(x): x * result in [1, 2)
    const int ex = ((__double2hiint(x) >> 20) & 0x7ff) - 1023;
    return __hiloint2double((1023 - ex) << 20, 0);
}

// sum_k a[k] * b[k] over one warp, fixed order: every warp that calls it with the same operands gets the same bits
__device__ __forceinline__ double warp_dot(const double* __restrict__ a, const double* __restrict__ b, int n) {
    const int lane = threadIdx.x & 31;
    double s0 = 0, s1 = 0;
    int k = lane;
    for (; k + 32 < n; k += 64) { s0 = fma(a[k], b[k], s0); s1 = fma(a[k + 32], b[k + 32], s1); }
    if (k < n) s0 = fma(a[k], b[k], s0);
    return warp_sum(s0 + s1);
}

// ---- S0 = XY_a^T XY_a with the pending deflation applied on the fly -------------------------------------------------------
__global__ void __launch_bounds__(256) wide_s0_kernel(WideArgs g) {
    const int K = g.K, M = g.M, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool first = g.comp == 0;
    double inv_tt = 0.0;
    if (!first) inv_tt = 1.0 / warp_dot(g.wprev, g.pprev, K);                     // tt^ = w^^T p^ of the previous component
    const int nent = M * (M + 1) / 2;
    for (int ent = blockIdx.x * 8 + wid; ent < nent; ent += gridDim.x * 8) {
        int i = 0, rem = ent;                                                      // entry -> (i, j), i <= j, row-major over the triangle
        while (rem >= M - i) { rem -= M - i; i++; }
        const int j = i + rem;
        const double* xi = g.XYold + (size_t)i * K;
        const double* xj = g.XYold + (size_t)j * K;
        const double ci = first ? 0.0 : g.qprev[i] * inv_tt, cj = first ? 0.0 : g.qprev[j] * inv_tt;
        double s0 = 0, s1 = 0;
        for (int k0 = lane; k0 < K; k0 += 64) {
            const int k1 = k0 + 32;
            const bool v1 = k1 < K;
            double a0 = xi[k0], b0 = xj[k0], a1 = v1 ? xi[k1] : 0.0, b1 = v1 ? xj[k1] : 0.0;
            if (!first) {
                const double p0 = g.pprev[k0], p1 = v1 ? g.pprev[k1] : 0.0;
                a0 = fma(-p0, ci, a0); b0 = fma(-p0, cj, b0);                      // pls.cpp:429
                a1 = fma(-p1, ci, a1); b1 = fma(-p1, cj, b1);
            }
            if (i == j) { g.XYnew[(size_t)i * K + k0] = a0; if (v1) g.XYnew[(size_t)i * K + k1] = a1; }
            s0 = fma(a0, b0, s0); s1 = fma(a1, b1, s1);
        }
        const double s = warp_sum(s0 + s1);
        if (lane == 0) { g.S0[(size_t)i * M + j] = s; g.S0[(size_t)j * M + i] = s; }
    }
    // W, P, Q of the previous component (pls.cpp:411, 427, 428): CTA 0
    if (!first && blockIdx.x == 0) {
        const double n = sqrt(warp_dot(g.wprev, g.wprev, K)), f = n * inv_tt;
        const int c = g.comp - 1;
        for (int k = threadIdx.x; k < K; k += 256) { g.W[(size_t)c * K + k] = g.wprev[k] / n; g.P[(size_t)c * K + k] = g.pprev[k] * f; }
        for (int m = threadIdx.x; m < M; m += 256) g.Q[(size_t)c * M + m] = g.qprev[m] * f;
    }
}

// outputs of the LAST component (the loop's s0 kernel emits component a - 1)
__global__ void __launch_bounds__(256) wide_emit_kernel(WideArgs g) {
    const int K = g.K, M = g.M;
    const double inv_tt = 1.0 / warp_dot(g.wprev, g.pprev, K);
    const double n = sqrt(warp_dot(g.wprev, g.wprev, K)), f = n * inv_tt;
    const int c = g.comp - 1;
    for (int k = threadIdx.x; k < K; k += 256) { g.W[(size_t)c * K + k] = g.wprev[k] / n; g.P[(size_t)c * K + k] = g.pprev[k] * f; }
    for (int m = threadIdx.x; m < M; m += 256) g.Q[(size_t)c * M + m] = g.qprev[m] * f;
}

// ---- dominant eigenvector of S0, w^ = XY q, q^ = XY^T w^ ------------------------------------------------------------------
__global__ void __launch_bounds__(WE_T, 1) wide_eig_kernel(WideArgs g) {
    extern __shared__ __align__(16) double sm[];
    const int K = g.K, M = g.M;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int gq = lane >> 2, qq = lane & 3;
    const int Mp = (M + 7) / 8 * 8, lds = Mp + 4, ntile = Mp / 8;
    const int ssz = Mp * lds;
    double* S0 = sm;
    double* Sa = S0 + ssz;
    double* Sb = Sa + ssz;
    double* dgA = Sb + ssz;           // Mp: compact diagonals of the eigen iterates (ping-pong)
    double* dgB = dgA + Mp;
    double* trs = dgB + Mp;           // 2
    double* wv = trs + 2;             // K
    __shared__ int s_flags[4];
    __shared__ int s_amax[2];
    __shared__ unsigned char pair_ta[136], pair_tb[136];
    const int npair = ntile * (ntile + 1) / 2;
    if (tid == 0) { int p = 0; for (int ta = 0; ta < ntile; ta++) for (int tb = ta; tb < ntile; tb++) { pair_ta[p] = (unsigned char)ta; pair_tb[p] = (unsigned char)tb; p++; } }
    for (int i = tid; i < 3 * ssz + 2 * Mp + 2; i += WE_T) sm[i] = 0.0;
    __syncthreads();
    for (int i = tid; i < M * M; i += WE_T) {
        const int r = i / M, c = i - r * M;
        const double v = g.S0[i];
        S0[r * lds + c] = v;
        if (r == c) dgA[r] = v;
    }
    __syncthreads();

    int bi = 0;
    bool degenerate = false;
    const double* src = S0;
    if (M != 1) {
        // projector onto the dominant eigenvector: B_{j+1} = (s_j B_j)^2, s_j a power of two; tr(B_{j+1}) / (s_j tr B_j)^2 -> 1
        // exactly when B_j has rank one. Trace and arg-max of the diagonal of iterate j are produced by the last working warp
        // while squaring j runs (pls_defl.cu, phase B).
        const int ppw = max(2, ((npair + WE_W - 2) / (WE_W - 1) + 1) / 2 * 2);   // pairs per warp (even), leaving the last warp for the trace when possible
        const int nwork = min(WE_W, (npair + ppw - 1) / ppw + 1);
        double T0 = 0;
        for (int a = lane; a < Mp; a += 32) T0 += dgA[a];
        T0 = warp_sum(T0);
        const bool deg0 = !(T0 > 0.0) || !(T0 < 1e300);
        degenerate = deg0;
        double* dst = Sa;
        const double* dgs = dgA;
        double* dgd = dgB;
        if (!degenerate && wid < nwork) {
            double u_prev = 0.0;
            for (int it = 0; it < 80; it++) {
                const double sc = (it == 0) ? pow2_inv_w(T0) : pow2_inv_w(u_prev * u_prev);
                const double sc2 = sc * sc;
                // A warp takes a run of consecutive tile pairs, two at a time: neighbours in the (ta, tb >= ta) list usually share
                // the row block ta, whose fragments are then loaded once for both (25 % fewer shared-memory fragment loads, the
                // bound of this loop at M ~ 50), and the two accumulator sets double the independent DMMA chains in flight.
                for (int p0 = wid * ppw; p0 < min(npair, (wid + 1) * ppw); p0 += 2) {
                    const bool two = p0 + 1 < min(npair, (wid + 1) * ppw);
                    const int ta0 = pair_ta[p0], tb0 = pair_tb[p0];
                    const int ta1 = two ? pair_ta[p0 + 1] : ta0, tb1 = two ? pair_tb[p0 + 1] : tb0;
                    const bool same_a = ta0 == ta1;
                    const double* pa0 = src + (ta0 * 8 + gq) * lds + qq;
                    const double* pb0 = src + (tb0 * 8 + gq) * lds + qq;   // B[k][n] = S[n][k] (symmetric)
                    const double* pa1 = src + (ta1 * 8 + gq) * lds + qq;
                    const double* pb1 = src + (tb1 * 8 + gq) * lds + qq;
                    double c[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}}, d[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
                    for (int ks = 0; ks < Mp; ks += 16) {
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            if (ks + 4 * u < Mp) {                                   // warp-uniform
                                const double a0 = pa0[ks + 4 * u], b0 = pb0[ks + 4 * u];
                                const double a1 = same_a ? a0 : pa1[ks + 4 * u], b1 = pb1[ks + 4 * u];
                                dmma884(c[u][0], c[u][1], a0, b0);
                                dmma884(d[u][0], d[u][1], a1, b1);
                            }
                        }
                    }
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        if (h == 1 && !two) break;
                        const int ta = h ? ta1 : ta0, tb = h ? tb1 : tb0;
                        const double (&x)[4][2] = h ? d : c;
                        const double c0 = ((x[0][0] + x[1][0]) + (x[2][0] + x[3][0])) * sc2, c1 = ((x[0][1] + x[1][1]) + (x[2][1] + x[3][1])) * sc2;
                        const int r = ta * 8 + gq, cc = tb * 8 + 2 * qq;
                        *(double2*)(dst + r * lds + cc) = make_double2(c0, c1);
                        if (ta != tb) { dst[cc * lds + r] = c0; dst[(cc + 1) * lds + r] = c1; }
                        else if ((gq >> 1) == qq) dgd[r] = (gq & 1) ? c1 : c0;
                    }
                }
                if (it > 0 && wid == nwork - 1) {   // tr(B_it) and the first arg-max of its diagonal
                    double t = 0, bv = -1.0; int bj = 0;
                    for (int a = lane; a < Mp; a += 32) { const double v = dgs[a]; t += v; if (a < M && v > bv) { bv = v; bj = a; } }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        t += __shfl_xor_sync(0xffffffffu, t, o);
                        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                        const int oi = __shfl_xor_sync(0xffffffffu, bj, o);
                        if (ov > bv || (ov == bv && oi < bj)) { bv = ov; bj = oi; }
                    }
                    if (lane == 0) { trs[it & 1] = t; s_amax[it & 1] = bj; }
                }
                asm volatile("bar.sync 1, %0;" ::"r"(nwork * 32) : "memory");
                bool conv = false;
                double u;
                if (it == 0) u = sc * T0;
                else {
                    const double Tj = trs[it & 1];
                    if (!(Tj > 0.0)) { degenerate = true; break; }
                    conv = Tj > (1.0 - PLS_EIG_DELTA) * u_prev * u_prev;
                    u = sc * Tj;
                    bi = s_amax[it & 1];
                }
                src = dst; dst = (dst == Sa) ? Sb : Sa;
                { const double* tswap = dgs; dgs = dgd; dgd = (double*)tswap; }
                if (conv) break;
                u_prev = u;
            }
            if (tid == 0) { s_flags[0] = degenerate ? 1 : 0; s_flags[1] = (src == Sa) ? 0 : (src == Sb ? 1 : 2); s_flags[2] = bi; }
        }
        __syncthreads();
        if (!deg0) {   // every warp adopts the outcome of the iteration
            degenerate = s_flags[0] != 0;
            src = (s_flags[1] == 0) ? Sa : (s_flags[1] == 1 ? Sb : S0);
            bi = s_flags[2];
        }
    }
    // w^ = XY q (pls.cpp:408), unnormalised; q = column bi of the projector (M == 1: w = XY, pls.cpp:403-404)
    const double* XY = g.XYnew;
    for (int k = tid; k < K; k += WE_T) {
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        if (M == 1 || degenerate) a0 = XY[k];
        else {
            const double* qc = src + bi * lds;
            int m = 0;
            for (; m + 3 < M; m += 4) {
                a0 = fma(XY[(size_t)m * K + k], qc[m], a0); a1 = fma(XY[(size_t)(m + 1) * K + k], qc[m + 1], a1);
                a2 = fma(XY[(size_t)(m + 2) * K + k], qc[m + 2], a2); a3 = fma(XY[(size_t)(m + 3) * K + k], qc[m + 3], a3);
            }
            for (; m < M; m++) a0 = fma(XY[(size_t)m * K + k], qc[m], a0);
        }
        const double v = (a0 + a1) + (a2 + a3);
        wv[k] = v;
        g.wcur[k] = v;
    }
    __syncthreads();
    // q^ = XY^T w^: one warp per response
    for (int m = wid; m < M; m += WE_W) {
        const double q = warp_dot(XY + (size_t)m * K, wv, K);
        if (lane == 0) g.qcur[m] = q;
    }
}

// ---- H <- H - (pending rank-one term), p^ = H w^ ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) wide_hw_kernel(WideArgs g) {
    const int K = g.K, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int i = blockIdx.x * 8 + wid;
    if (i >= K) return;
    const bool first = g.comp == 0;
    double inv_tt = 0.0, pi = 0.0;
    if (!first) { inv_tt = 1.0 / warp_dot(g.wprev, g.pprev, K); pi = g.pprev[i]; }
    double* hrow = g.H + (size_t)i * K;
    double s0 = 0, s1 = 0;
    for (int j0 = lane; j0 < K; j0 += 64) {
        const int j1 = j0 + 32;
        const bool v1 = j1 < K;
        double x0 = hrow[j0], x1 = v1 ? hrow[j1] : 0.0;
        if (!first) {
            x0 = fma(-(pi * g.pprev[j0]), inv_tt, x0);                 // p_i p_j is the same product in row i and row j: H stays symmetric
            hrow[j0] = x0;
            if (v1) { x1 = fma(-(pi * g.pprev[j1]), inv_tt, x1); hrow[j1] = x1; }
        }
        s0 = fma(x0, g.wcur[j0], s0);
        if (v1) s1 = fma(x1, g.wcur[j1], s1);
    }
    const double p = warp_sum(s0 + s1);
    if (lane == 0) g.pcur[i] = p;
}

size_t wide_eig_smem(int K, int M) {
    const size_t Mp = (size_t)(M + 7) / 8 * 8;
    return sizeof(double) * (3 * Mp * (Mp + 4) + 2 * Mp + 2 + (size_t)K + 8);
}

}  // namespace

bool pls_wide_fits(const abcb200_ctx* ctx, int K, int M) { return M <= 128 && wide_eig_smem(K, M) + 1024 <= (size_t)ctx->smem_optin; }

size_t pls_wide_ws_bytes(int K, int M, int A) {
    return align_up((size_t)K * K * 8, 256) + 2 * align_up((size_t)K * M * 8, 256) + 4 * align_up((size_t)K * 8, 256) + 2 * align_up((size_t)M * 8, 256) +
           align_up((size_t)M * M * 8, 256) + align_up((size_t)A * A * 8, 256) + 4096;
}

// Buffers of one fit (H = X^T X deflated in place, ping-pong XY / w^ / p^ / q^): allocated from the context's workspace.
int pls_wide_begin(abcb200_ctx* ctx, const double* XX, const double* XY, const PlsFactors& f, WideJob* job) {
    const int K = f.K, M = f.M, A = f.A;
    WideJob& j = *job;
    j.K = K; j.M = M; j.A = A; j.XY0 = XY; j.W = f.W; j.P = f.P; j.Q = f.Q;
    j.H = ws_new<double>(ctx, (size_t)K * K);
    for (int s = 0; s < 2; s++) { j.XYb[s] = ws_new<double>(ctx, (size_t)K * M); j.wb[s] = ws_new<double>(ctx, K); j.pb[s] = ws_new<double>(ctx, K); j.qb[s] = ws_new<double>(ctx, M); }
    j.S0 = ws_new<double>(ctx, (size_t)M * M);
    if (!j.H || !j.XYb[0] || !j.XYb[1] || !j.wb[0] || !j.wb[1] || !j.pb[0] || !j.pb[1] || !j.qb[0] || !j.qb[1] || !j.S0)
        ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in pls_wide");
    CUDA_TRY(ctx, cudaMemcpyAsync(j.H, XX, sizeof(double) * K * K, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaFuncSetAttribute(wide_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_eig_smem(K, M)));
    return ABCB200_OK;
}

// Components [c0, c1): three launches each; W, P, Q of every component of the range are in place when the launches finish.
int pls_wide_block(abcb200_ctx* ctx, const WideJob* job, int c0, int c1) {
    const WideJob& j = *job;
    const int K = j.K, M = j.M;
    const size_t smem = wide_eig_smem(K, M);
    const int nent = M * (M + 1) / 2;
    const int s0_grid = std::max(1, std::min((nent + 7) / 8, 4 * ctx->sm_count));
    WideArgs g;
    g.K = K; g.M = M; g.A = j.A; g.H = j.H; g.S0 = j.S0; g.W = j.W; g.P = j.P; g.Q = j.Q;
    for (int a = c0; a < c1; a++) {
        const int cur = a & 1, prv = cur ^ 1;
        g.comp = a;
        g.XYold = (a == 0) ? j.XY0 : j.XYb[prv]; g.XYnew = j.XYb[cur];
        g.wprev = j.wb[prv]; g.pprev = j.pb[prv]; g.qprev = j.qb[prv];
        g.wcur = j.wb[cur]; g.pcur = j.pb[cur]; g.qcur = j.qb[cur];
        LAUNCH(ctx, wide_s0_kernel, s0_grid, 256, 0, g);
        LAUNCH(ctx, wide_eig_kernel, 1, WE_T, smem, g);
        LAUNCH(ctx, wide_hw_kernel, (K + 7) / 8, 256, 0, g);
    }
    {   // W, P, Q of the last component of the range (the next component's launches would write the same values again)
        const int last = (c1 - 1) & 1;
        g.comp = c1;
        g.wprev = j.wb[last]; g.pprev = j.pb[last]; g.qprev = j.qb[last];
        LAUNCH(ctx, wide_emit_kernel, 1, 256, 0, g);
    }
    return ABCB200_OK;
}

// Component loop from XX (K x K) and XY (K x M, ld K): fills W, P, Q and R. Three launches per component.
int pls_wide_dev(abcb200_ctx* ctx, const double* XX, const double* XY, const PlsFactors& f) {
    WideJob job;
    ABC_TRY(pls_wide_begin(ctx, XX, XY, f, &job));
    double* U = ws_new<double>(ctx, (size_t)f.A * f.A);
    if (!U) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in pls_wide");
    kernel_begin(ctx, 0);
    ABC_TRY(pls_wide_block(ctx, &job, 0, f.A));
    kernel_end(ctx, 0);
    return pls_ur_dev(ctx, f, U);
}
