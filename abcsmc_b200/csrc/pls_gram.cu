// pls_gram.cu — the component loop of kernel PLS as ONE persistent CTA.
//
// Reference: PLS::Model::plsr, lib/PLS/src/pls.cpp:400-435. After XY = X^T Y (:396) and XX = X^T X (:398) every
// quantity of the loop is a function of those two small matrices: t = X r only enters through tt = t^T t = r^T XX r and
// p = X^T t / tt = XX r / tt (:418-427), which is how the reference's own KERNEL_TYPE2 evaluates them (:422-424). So the
// N-long passes of KERNEL_TYPE1 collapse into the two Gram products (one read of X) and A components of O(K^2 + K*M + M^3)
// on-chip work. That work is a strict dependency chain (each component needs the deflated XY of the previous one), so
// it is latency bound: the design goal is the fewest, shortest barrier-to-barrier phases, not throughput.
//
//   phase A  S0 = XY^T XY (M x M)                         DMMA 8x8 tiles of the upper triangle, mirrored
//   phase B  dominant eigenvector q of S0                 B <- (B / tr B)^2 by DMMA; tr(B_next) = sum l_i^2 / (sum l_i)^2
//                                                         -> 1 exactly when B is rank one: the convergence test is free
//   phase C  w = XY q / |XY q|                            (pls.cpp:408-411)
//   phase D  c = P^T w ; r = w - R c                      (pls.cpp:412-416)
//   phase E  p = XX r / tt ; q = XY^T r / tt ; XY -= tt p q^T   (pls.cpp:422-429)
//
// XY (and XX, P, R when they fit) live in shared memory with a leading dimension = 4 mod 16 doubles, which makes the
// DMMA fragment loads and the 4-lanes-per-row matrix-vector products bank-conflict free; otherwise they stay in
// global memory (L2 resident: one CTA touches them).
#include <stdlib.h>

#include "kernels.cuh"

namespace {

constexpr int GT = 512;             // threads of the persistent CTA
constexpr int GW = GT / 32;

struct GramArgs {
    const double* XX;   // K x K (ld K), symmetric
    const double* XY0;  // K x M (ld K)
    double* XYg;        // K x M global scratch (used when XY does not fit in shared memory)
    double *W, *P, *R, *Q;
    double* Rt;         // K x A row-major copy of R (row k = R[k, :]) so that r = w - R c is a row matvec
    long long* prof;    // optional per-phase clock totals (debug), 8 entries
    int K, M, A;
    int ldk;            // leading dimension of the shared-memory K-vectors' matrices
    int lda;            // leading dimension of the shared-memory copy of Rt
    int xy_smem, xx_smem, pr_smem;
};

__device__ __forceinline__ double warp_trace_g(const double* S, int M, int lds) {
    double t = 0;
    for (int a = threadIdx.x & 31; a < M; a += 32) t += S[a * lds + a];
    return warp_sum(t);
}

// two-level block sum with one barrier pair: every thread returns the total (fixed order -> deterministic)
__device__ __forceinline__ double block_sum_g(double v, double* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double s = 0;
#pragma unroll
    for (int w = 0; w < GW; w++) s += red[w];
    return s;
}

// y[row] = sum_k Mat[row*ld + k] * x[k] for row < nrows, L lanes per row (k interleaved by L), result in lane part 0.
// Mat may be shared or global memory (generic pointer); x is in shared memory. Loads are issued in batches of 8
// independent requests per thread (branch-free, clamped addresses) so that an L2-resident matrix is read at
// bandwidth instead of one latency per element.
template <int L, typename F>
__device__ __forceinline__ void matvec_l(const double* __restrict__ Mat, int ld, int nrows, int ncols, const double* x, F&& sink) {
    const int part = threadIdx.x % L, rloc = threadIdx.x / L;
    for (int row0 = 0; row0 < nrows; row0 += GT / L) {
        const int row = row0 + rloc;
        const bool rv = row < nrows;
        const double* m = Mat + (size_t)(rv ? row : 0) * ld;
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        for (int k = part; k < ncols; k += 8 * L) {
            double v[8], xv[8];
#pragma unroll
            for (int u = 0; u < 8; u++) { const int kk = k + u * L; const bool ok = kk < ncols; v[u] = m[ok ? kk : 0]; xv[u] = ok ? x[kk] : 0.0; }
            a0 = fma(v[0], xv[0], a0); a1 = fma(v[1], xv[1], a1); a2 = fma(v[2], xv[2], a2); a3 = fma(v[3], xv[3], a3);
            a0 = fma(v[4], xv[4], a0); a1 = fma(v[5], xv[5], a1); a2 = fma(v[6], xv[6], a2); a3 = fma(v[7], xv[7], a3);
        }
        double a = (a0 + a1) + (a2 + a3);
#pragma unroll
        for (int o = 1; o < L; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (part == 0 && rv) sink(row, a);
    }
}
// lanes per row so that one pass of the CTA covers `nrows` rows when possible (fewest serial batches per thread)
template <typename F>
__device__ __forceinline__ void matvec(const double* Mat, int ld, int nrows, int ncols, const double* x, F&& sink) {
    if (nrows * 16 <= GT) matvec_l<16>(Mat, ld, nrows, ncols, x, sink);
    else if (nrows * 8 <= GT) matvec_l<8>(Mat, ld, nrows, ncols, x, sink);
    else if (nrows * 4 <= GT) matvec_l<4>(Mat, ld, nrows, ncols, x, sink);
    else if (nrows * 2 <= GT) matvec_l<2>(Mat, ld, nrows, ncols, x, sink);
    else matvec_l<1>(Mat, ld, nrows, ncols, x, sink);
}

__global__ void __launch_bounds__(GT, 1) pls_gram_kernel(GramArgs g) {
    extern __shared__ __align__(16) double sm[];
    const int K = g.K, M = g.M, A = g.A, ldk = g.ldk;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int gq = lane >> 2, qq = lane & 3;
    const int Mp = (M + 7) / 8 * 8, lds = Mp + 4, ntile = Mp / 8;
    const int ssz = Mp * lds;
    double* S0 = sm;
    double* Sa = S0 + ssz;
    double* Sb = Sa + ssz;
    double* qv = Sb + ssz;            // Mp
    double* wv = qv + Mp;             // K
    double* rv = wv + K;              // K
    double* pv = rv + K;              // K
    double* cv = pv + K;              // A
    double* red = cv + A;             // 32
    double* dyn = red + 32;
    double* XY = g.XYg;
    int ldxy = K;
    if (g.xy_smem) { XY = dyn; ldxy = ldk; dyn += (size_t)M * ldk; }
    const double* XX = g.XX;
    int ldxx = K;
    if (g.xx_smem) { double* XXs = dyn; dyn += (size_t)K * ldk; for (int i = tid; i < K * K; i += GT) { const int b = i / K, k = i - b * K; XXs[(size_t)b * ldk + k] = g.XX[i]; } XX = XXs; ldxx = ldk; }
    double* Ps = g.P; double* Rts = g.Rt;
    int ldpr = K, ldrt = A;
    if (g.pr_smem) { Ps = dyn; dyn += (size_t)A * ldk; Rts = dyn; dyn += (size_t)K * g.lda; ldpr = ldk; ldrt = g.lda; }
    long long tprev = clock64();
#define PROF(slot) do { if (g.prof && tid == 0) { const long long tn = clock64(); g.prof[slot] += tn - tprev; tprev = tn; } } while (0)
    for (int i = tid; i < K * M; i += GT) { const int m = i / K, k = i - m * K; XY[(size_t)m * ldxy + k] = g.XY0[i]; }
    for (int i = tid; i < 3 * ssz; i += GT) S0[i] = 0.0;      // zero padding of the M x M work matrices
    __syncthreads();

    for (int comp = 0; comp < A; comp++) {
        double wk = 0.0;                                      // thread k < K owns w[k] (K <= GT fast path; else loop)
        if (M == 1) {                                                                           // pls.cpp:403-404
            for (int k = tid; k < K; k += GT) wv[k] = XY[k];
        } else {
            // ---- phase A: S0 = XY^T XY (pls.cpp:406), upper-triangle tiles mirrored (bitwise symmetric) ---------
            const int npair = ntile * (ntile + 1) / 2;
            for (int pidx = wid; pidx < npair; pidx += GW) {
                int ta = 0, rem = pidx;
                while (rem >= ntile - ta) { rem -= ntile - ta; ta++; }
                const int tb = ta + rem;
                const int ca = ta * 8 + gq, cb = tb * 8 + gq;
                const bool va = ca < M, vb = cb < M;
                const double* pa = XY + (size_t)min(ca, M - 1) * ldxy;
                const double* pb = XY + (size_t)min(cb, M - 1) * ldxy;
                double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;   // two accumulator pairs: halves the dependent DMMA chain
                int k0 = 0;
                for (; k0 + 8 <= K; k0 += 8) {
                    const double a0 = pa[k0 + qq], b0 = pb[k0 + qq], a1 = pa[k0 + 4 + qq], b1 = pb[k0 + 4 + qq];
                    dmma884(c0, c1, va ? a0 : 0.0, vb ? b0 : 0.0);
                    dmma884(d0, d1, va ? a1 : 0.0, vb ? b1 : 0.0);
                }
                for (; k0 < K; k0 += 4) {
                    const int k = k0 + qq;
                    const bool kv = k < K;
                    const double av = pa[kv ? k : K - 1], bv = pb[kv ? k : K - 1];
                    dmma884(c0, c1, (va && kv) ? av : 0.0, (vb && kv) ? bv : 0.0);
                }
                c0 += d0; c1 += d1;
                const int r = ta * 8 + gq, c = tb * 8 + 2 * qq;
                S0[r * lds + c] = c0; S0[r * lds + c + 1] = c1;
                if (ta != tb) { S0[c * lds + r] = c0; S0[(c + 1) * lds + r] = c1; }
            }
            __syncthreads();
            PROF(0);
            // ---- phase B: dominant eigenvector by trace-normalised repeated squaring ---------------------------
            const double* src = S0;
            double* dst = Sa;
            double tr = warp_trace_g(src, M, lds);
            bool degenerate = !(tr > 0.0);                     // zero or NaN matrix (warp- and block-uniform)
            if (!degenerate) {
                for (int it = 0; it < 80; it++) {
                    const double inv = 1.0 / tr, inv2 = inv * inv;
                    for (int tile = wid; tile < ntile * ntile; tile += GW) {
                        const int ta = tile / ntile, tb = tile - ta * ntile;
                        const double* pa = src + (ta * 8 + gq) * lds + qq;
                        const double* pb = src + (tb * 8 + gq) * lds + qq;   // B[k][n] = S[n][k] (symmetric)
                        double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
                        int ks = 0;
                        for (; ks + 8 <= Mp; ks += 8) { dmma884(c0, c1, pa[ks], pb[ks]); dmma884(d0, d1, pa[ks + 4], pb[ks + 4]); }
                        if (ks < Mp) dmma884(c0, c1, pa[ks], pb[ks]);
                        double* po = dst + (ta * 8 + gq) * lds + tb * 8 + 2 * qq;
                        po[0] = (c0 + d0) * inv2; po[1] = (c1 + d1) * inv2;
                    }
                    __syncthreads();
                    tr = warp_trace_g(dst, M, lds);            // = sum l_i^2 / (sum l_i)^2 of the previous iterate
                    src = dst; dst = (dst == Sa) ? Sb : Sa;
                    if (1.0 - tr < 1e-9) break;                // previous iterate had l2/l1 < ~5e-10: this one is rank one to 1e-18
                    if (!(tr > 0.0)) { degenerate = true; break; }
                }
            }
            if (wid == 0) {   // q: normalised column (largest diagonal) of the projector, one power step with S0, renormalised
                if (degenerate) {
                    for (int a = lane; a < Mp; a += 32) qv[a] = (a == 0) ? 1.0 : 0.0;
                } else {
                    double bv = -1.0; int bi = 0;
                    for (int a = lane; a < M; a += 32) { const double v = src[a * lds + a]; if (v > bv) { bv = v; bi = a; } }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                    }
                    for (int a = lane; a < Mp; a += 32) qv[a] = (a < M) ? src[a * lds + bi] : 0.0;
                    __syncwarp();
                    double v[4] = {0, 0, 0, 0};               // M <= 128
                    double nn = 0;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int a = lane + 32 * j;
                        if (a < M) {
                            double acc0 = 0, acc1 = 0;
                            int l = 0;
                            for (; l + 1 < M; l += 2) { acc0 = fma(S0[l * lds + a], qv[l], acc0); acc1 = fma(S0[(l + 1) * lds + a], qv[l + 1], acc1); }
                            if (l < M) acc0 = fma(S0[l * lds + a], qv[l], acc0);
                            v[j] = acc0 + acc1;
                            nn = fma(v[j], v[j], nn);
                        }
                    }
                    nn = warp_sum(nn);
                    const double sc = 1.0 / sqrt(nn);
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 4; j++) { const int a = lane + 32 * j; if (a < M) qv[a] = v[j] * sc; }
                }
            }
            __syncthreads();
            PROF(1);
            // ---- phase C: w = XY q (pls.cpp:408) --------------------------------------------------------------------
            for (int k = tid; k < K; k += GT) {
                double a0 = 0, a1 = 0;
                int m = 0;
                for (; m + 1 < M; m += 2) { a0 = fma(XY[(size_t)m * ldxy + k], qv[m], a0); a1 = fma(XY[(size_t)(m + 1) * ldxy + k], qv[m + 1], a1); }
                if (m < M) a0 = fma(XY[(size_t)m * ldxy + k], qv[m], a0);
                wv[k] = a0 + a1;
            }
        }
        // every thread only re-reads the wv entries it wrote itself until the barrier inside block_sum_g
        double ww = 0;
        for (int k = tid; k < K; k += GT) { const double v = wv[k]; ww = fma(v, v, ww); }
        ww = block_sum_g(ww, red);
        const double wn = sqrt(ww);
        for (int k = tid; k < K; k += GT) { wk = wv[k] / wn; wv[k] = wk; g.W[(size_t)comp * K + k] = wk; }    // pls.cpp:411
        __syncthreads();
        PROF(2);
        // ---- phase D: c_j = P_j^T w (pls.cpp:415); r = w - sum_j c_j R_j -------------------------------------------
        matvec(Ps, ldpr, comp, K, wv, [&](int j, double v) { cv[j] = v; });
        __syncthreads();
        matvec(Rts, ldrt, K, comp, cv, [&](int k, double v) {
            const double r = wv[k] - v;
            rv[k] = r;
            g.R[(size_t)comp * K + k] = r;
            Rts[(size_t)k * ldrt + comp] = r;
        });
        __syncthreads();
        PROF(3);
        // ---- phase E: p = XX r, tt = r^T XX r (pls.cpp:422-424); q = XY^T r (pls.cpp:428) ------------------------------
        matvec(XX, ldxx, K, K, rv, [&](int b, double v) { pv[b] = v; });
        matvec(XY, ldxy, M, K, rv, [&](int m, double v) { qv[m] = v; });
        __syncthreads();
        double t = 0;
        for (int k = tid; k < K; k += GT) t = fma(pv[k], rv[k], t);
        const double tt = block_sum_g(t, red);
        for (int k = tid; k < K; k += GT) {                                                    // pls.cpp:427
            const double p = pv[k] / tt;
            pv[k] = p;
            g.P[(size_t)comp * K + k] = p;
            if (g.pr_smem) Ps[(size_t)comp * ldpr + k] = p;   // (global P doubles as Ps otherwise)
        }
        for (int m = tid; m < M; m += GT) { const double q = qv[m] / tt; qv[m] = q; g.Q[(size_t)comp * M + m] = q; }
        __syncthreads();
        for (int i = tid; i < K * M; i += GT) {                                                // pls.cpp:429
            const int m = i / K, k = i - m * K;
            XY[(size_t)m * ldxy + k] -= (pv[k] * qv[m]) * tt;
        }
        __syncthreads();
        PROF(4);
    }
#undef PROF
}

}  // namespace

size_t pls_gram_ws_bytes(const abcb200_ctx* ctx, int64_t n, int K, int M) {
    return 2 * align_up((size_t)K * M * 8, 256) + align_up((size_t)K * K * 8, 256) + 512 + align_up((size_t)K * K * 8, 256) + atb_ws_bytes(ctx, n, K, M) + atb_ws_bytes(ctx, n, K, K) + 1024;
}

// Fits f.A components from X (n x K), Y (n x M): two Gram products (DMMA) + the persistent component-loop CTA.
// Fills W, P, R, Q; T is left to the caller (T = X R, pls.cpp:418 — launch_xb).
int pls_fit_gram_dev(abcb200_ctx* ctx, const double* X, int64_t ldx, const double* Y, int64_t ldy, const PlsFactors& f) {
    const int K = f.K, M = f.M, A = f.A;
    const int64_t n = f.n;
    if (M > 128) ABC_FAIL(ctx, ABCB200_EINVAL, "pls_fit: M=%d responses exceed the on-chip eigen-solver limit (128)", M);
    double* XY = ws_new<double>(ctx, (size_t)K * M);
    double* XYg = ws_new<double>(ctx, (size_t)K * M);
    double* XX = ws_new<double>(ctx, (size_t)K * K);
    double* Rt = ws_new<double>(ctx, (size_t)K * A);
    long long* prof = ws_new<long long>(ctx, 8);
    if (!XY || !XYg || !XX || !Rt || !prof) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in pls_fit_gram");
    ABC_TRY(launch_atb(ctx, X, ldx, K, Y, ldy, M, n, XY));        // pls.cpp:396
    ABC_TRY(launch_atb(ctx, X, ldx, K, X, ldx, K, n, XX));        // pls.cpp:398
    GramArgs g;
    g.XX = XX; g.XY0 = XY; g.XYg = XYg; g.W = f.W; g.P = f.P; g.R = f.R; g.Q = f.Q; g.Rt = Rt; g.K = K; g.M = M; g.A = A;
    int ldk = K;
    while (ldk % 16 != 4) ldk++;
    g.ldk = ldk;
    int lda = A;
    while (lda % 16 != 4) lda++;
    g.lda = lda;
    static const bool want_prof = getenv("ABCB200_PLS_PROF") != nullptr;     // debug: per-phase clock totals on stderr
    g.prof = want_prof ? prof : nullptr;
    if (want_prof) CUDA_TRY(ctx, cudaMemsetAsync(prof, 0, 8 * sizeof(long long), ctx->stream));
    const size_t Mp = (size_t)(M + 7) / 8 * 8;
    const size_t fixed = sizeof(double) * (3 * Mp * (Mp + 4) + Mp + 3 * (size_t)K + A + 32) + 256;
    const size_t budget = (size_t)ctx->smem_optin;
    if (fixed > budget) ABC_FAIL(ctx, ABCB200_EINVAL, "pls_fit: K=%d M=%d A=%d need %zu B of shared memory (> %zu)", K, M, A, fixed, budget);
    size_t used = fixed;
    const size_t xy_b = (size_t)M * ldk * 8, xx_b = (size_t)K * ldk * 8, pr_b = ((size_t)A * ldk + (size_t)K * lda) * 8;
    g.xy_smem = (used + xy_b <= budget) ? 1 : 0; if (g.xy_smem) used += xy_b;
    g.xx_smem = (used + xx_b <= budget) ? 1 : 0; if (g.xx_smem) used += xx_b;
    g.pr_smem = (used + pr_b <= budget) ? 1 : 0; if (g.pr_smem) used += pr_b;
    CUDA_TRY(ctx, cudaFuncSetAttribute(pls_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)used));
    LAUNCH(ctx, pls_gram_kernel, 1, GT, used, g);
    if (want_prof) {
        long long h[8];
        CUDA_TRY(ctx, cudaMemcpyAsync(h, prof, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        fprintf(stderr, "[pls_gram K=%d M=%d A=%d smem xy/xx/pr=%d/%d/%d] cycles per component: S0 %.0f | eigen %.0f | w %.0f | c,r %.0f | p,q,deflate %.0f\n", K, M, A,
                g.xy_smem, g.xx_smem, g.pr_smem, (double)h[0] / A, (double)h[1] / A, (double)h[2] / A, (double)h[3] / A, (double)h[4] / A);
    }
    return ABCB200_OK;
}
