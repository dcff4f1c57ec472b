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

// y[row] = sum_k Mat[row*ld + k] * x[k*xs] for row < nrows, L lanes per row (k interleaved by L), result in lane part 0.
// Mat may be shared or global memory (generic pointer); x is in shared memory. A thread issues up to MV_B independent
// loads (branch-free, clamped addresses) before the first use, so an L2-resident matrix costs about one L2 latency per
// pass of the CTA over the rows instead of one per element.
constexpr int MV_B = 20;
template <int L, int B, typename F>
__device__ __forceinline__ void matvec_l(const double* __restrict__ Mat, int ld, int nrows, int ncols, const double* x, int xs, F&& sink) {
    const int part = threadIdx.x % L, rloc = threadIdx.x / L;
    for (int row0 = 0; row0 < nrows; row0 += GT / L) {
        const int row = row0 + rloc;
        const bool rv = row < nrows;
        const double* m = Mat + (size_t)(rv ? row : 0) * ld;
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        for (int k = part; k < ncols; k += B * L) {
            double v[B];
#pragma unroll
            for (int u = 0; u < B; u++) { const int kk = k + u * L; v[u] = m[kk < ncols ? kk : 0]; }
#pragma unroll
            for (int u = 0; u < B; u += 4) {
                const int k0 = k + u * L, k1 = k0 + L, k2 = k1 + L, k3 = k2 + L;
                a0 = fma(v[u], k0 < ncols ? x[(size_t)k0 * xs] : 0.0, a0);
                a1 = fma(v[u + 1], k1 < ncols ? x[(size_t)k1 * xs] : 0.0, a1);
                a2 = fma(v[u + 2], k2 < ncols ? x[(size_t)k2 * xs] : 0.0, a2);
                a3 = fma(v[u + 3], k3 < ncols ? x[(size_t)k3 * xs] : 0.0, a3);
            }
        }
        double a = (a0 + a1) + (a2 + a3);
#pragma unroll
        for (int o = 1; o < L; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (part == 0 && rv) sink(row, a);
    }
}
// Lanes per row: enough that a thread's share of a row fits one batch of loads (fewest L2 round trips), but no more
// than needed to cover the rows in one pass of the CTA when the rows are short.
template <typename F>
__device__ __forceinline__ void matvec(const double* Mat, int ld, int nrows, int ncols, const double* x, int xs, F&& sink) {
    int L = 1;
    while (L < 16 && (ncols + L - 1) / L > MV_B) L <<= 1;         // one batch per row share
    while (L < 16 && nrows * (2 * L) <= GT) L <<= 1;              // idle threads: split the rows further
    const bool small = (ncols + L - 1) / L <= 4;                  // short row shares: a 4-deep batch, no wasted slots
    if (small) {
        if (L == 16) matvec_l<16, 4>(Mat, ld, nrows, ncols, x, xs, sink);
        else if (L == 8) matvec_l<8, 4>(Mat, ld, nrows, ncols, x, xs, sink);
        else if (L == 4) matvec_l<4, 4>(Mat, ld, nrows, ncols, x, xs, sink);
        else if (L == 2) matvec_l<2, 4>(Mat, ld, nrows, ncols, x, xs, sink);
        else matvec_l<1, 4>(Mat, ld, nrows, ncols, x, xs, sink);
    } else {
        if (L == 16) matvec_l<16, MV_B>(Mat, ld, nrows, ncols, x, xs, sink);
        else if (L == 8) matvec_l<8, MV_B>(Mat, ld, nrows, ncols, x, xs, sink);
        else if (L == 4) matvec_l<4, MV_B>(Mat, ld, nrows, ncols, x, xs, sink);
        else if (L == 2) matvec_l<2, MV_B>(Mat, ld, nrows, ncols, x, xs, sink);
        else matvec_l<1, MV_B>(Mat, ld, nrows, ncols, x, xs, sink);
    }
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
    double* red = cv + A;             // 32: [0,16) |w|^2 partials, [16,32) tt partials
    double* dgA = red + 32;           // Mp: compact diagonals of the eigen iterates (ping-pong)
    double* dgB = dgA + Mp;           // Mp
    double* trs = dgB + Mp;           // 2: traces of the eigen iterates (ping-pong)
    double* dyn = trs + 2;
    __shared__ int s_flags[4];
    __shared__ unsigned char pair_ta[136], pair_tb[136];      // upper-triangle tile pairs, ntile <= 16
    const int npair = ntile * (ntile + 1) / 2;
    if (tid == 0) { int p = 0; for (int ta = 0; ta < ntile; ta++) for (int tb = ta; tb < ntile; tb++) { pair_ta[p] = (unsigned char)ta; pair_tb[p] = (unsigned char)tb; p++; } }
    double* XY = g.XYg;
    int ldxy = K;
    if (g.xy_smem) { XY = dyn; ldxy = ldk; dyn += (size_t)M * ldk; }
    const double* XX = g.XX;
    int ldxx = K;
    if (g.xx_smem) { double* XXs = dyn; dyn += (size_t)K * ldk; for (int i = tid; i < K * K; i += GT) { const int b = i / K, k = i - b * K; XXs[(size_t)b * ldk + k] = g.XX[i]; } XX = XXs; ldxx = ldk; }
    double* Ps = g.P; double* Rts = g.Rt;
    int ldpr = K, ldrt = A;
    if (g.pr_smem) { Ps = dyn; dyn += (size_t)A * ldk; Rts = dyn; dyn += (size_t)K * g.lda; ldpr = ldk; ldrt = g.lda; }
    long long tprev = clock64(), pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // per-phase clock totals of thread 0, kept in registers
#define PROF(slot) do { if (g.prof && tid == 0) { const long long tn = clock64(); pacc[slot] += tn - tprev; tprev = tn; } } while (0)
    for (int i = tid; i < K * M; i += GT) { const int m = i / K, k = i - m * K; XY[(size_t)m * ldxy + k] = g.XY0[i]; }
    for (int i = tid; i < 3 * ssz; i += GT) S0[i] = 0.0;      // zero padding of the M x M work matrices
    __syncthreads();

    for (int comp = 0; comp < A; comp++) {
        if (M != 1) {
            // ---- phase A: S0 = XY^T XY (pls.cpp:406), upper-triangle tiles mirrored (bitwise symmetric) ---------
            for (int pidx = wid; pidx < npair; pidx += GW) {
                const int ta = pair_ta[pidx], tb = pair_tb[pidx];
                const int ca = ta * 8 + gq, cb = tb * 8 + gq;
                const bool va = ca < M, vb = cb < M;
                const double* pa = XY + (size_t)min(ca, M - 1) * ldxy;
                const double* pb = XY + (size_t)min(cb, M - 1) * ldxy;
                double c[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};   // four accumulator pairs: short dependent DMMA chains
                int k0 = 0;
                for (; k0 + 16 <= K; k0 += 16) {
                    double av[4], bv[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) { av[u] = pa[k0 + 4 * u + qq]; bv[u] = pb[k0 + 4 * u + qq]; }
#pragma unroll
                    for (int u = 0; u < 4; u++) dmma884(c[u][0], c[u][1], va ? av[u] : 0.0, vb ? bv[u] : 0.0);
                }
                if (k0 < K) {   // up to four ragged steps, predicated (register arrays stay statically indexed)
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int k = k0 + 4 * u + qq;
                        const bool kv = k < K;
                        const double av = pa[kv ? k : K - 1], bv = pb[kv ? k : K - 1];
                        if (k0 + 4 * u < K) dmma884(c[u][0], c[u][1], (va && kv) ? av : 0.0, (vb && kv) ? bv : 0.0);
                    }
                }
                const double c0 = (c[0][0] + c[1][0]) + (c[2][0] + c[3][0]), c1 = (c[0][1] + c[1][1]) + (c[2][1] + c[3][1]);
                const int r = ta * 8 + gq, cc = tb * 8 + 2 * qq;
                S0[r * lds + cc] = c0; S0[r * lds + cc + 1] = c1;
                if (ta != tb) { S0[cc * lds + r] = c0; S0[(cc + 1) * lds + r] = c1; }
                else if ((gq >> 1) == qq) dgA[r] = (gq & 1) ? c1 : c0;        // compact copy of the diagonal
            }
            __syncthreads();
            PROF(0);
            // ---- phase B: dominant eigenvector by trace-normalised repeated squaring ---------------------------
            // B_{j+1} = (s_j B_j)^2 with s_j a power of two (exact scaling, no division). With u_j = s_j tr(B_j):
            // tr(B_{j+1}) / u_j^2 = sum l_i^2 / (sum l_i)^2 -> 1 exactly when B_j has rank one, so
            // "tr(B_{j+1}) > (1 - d) u_j^2" says B_j had l2/l1 < d/2; noticed one squaring late, the iterate used is B_{j+2}:
            // ratios (d/2)^4 (PLS_EIG_DELTA, kernels.cuh).
            // The trace of an iterate is summed by one otherwise idle warp WHILE the next squaring runs (it is never on the
            // critical path): the scale of step j comes from the bound tr(B_j) <= u_{j-1}^2 (within a factor M of the
            // truth, re-centred every step) and convergence is noticed one squaring late, which costs nothing in accuracy.
            auto warp_diag = [&](const double* dg) {
                double t = 0;
                for (int a = lane; a < Mp; a += 32) t += dg[a];
                return warp_sum(t);
            };
            auto pow2_inv = [](double x) {     // 2^-exponent(x): x * result in [1, 2)
                const int ex = ((__double2hiint(x) >> 20) & 0x7ff) - 1023;
                return __hiloint2double((1023 - ex) << 20, 0);
            };
            const double* src = S0;
            double* dst = Sa;
            const double* dgs = dgA;
            double* dgd = dgB;
            const double T0 = warp_diag(dgs);                 // same bits in every warp
            const bool deg0 = !(T0 > 0.0) || !(T0 < 1e300);   // zero, NaN or inf matrix (uniform over the CTA)
            bool degenerate = deg0;
            const double sc0 = deg0 ? 1.0 : pow2_inv(T0);
            // Only the warps that own tile pairs (plus the trace warp) take part in the iteration and meet at a named barrier
            // sized for them: a CTA-wide barrier costs several hundred cycles here (it drains every warp's pending stores).
            const int nwork = min(GW, npair + 1);
            if (!degenerate && wid < nwork) {
                double u_prev = 0.0;
                for (int it = 0; it < 80; it++) {
                    const double sc = (it == 0) ? pow2_inv(T0) : pow2_inv(u_prev * u_prev);
                    const double sc2 = sc * sc;
                    for (int pidx = wid; pidx < npair; pidx += GW) {
                        const int ta = pair_ta[pidx], tb = pair_tb[pidx];
                        const double* pa = src + (ta * 8 + gq) * lds + qq;
                        const double* pb = src + (tb * 8 + gq) * lds + qq;   // B[k][n] = S[n][k] (symmetric)
                        double c[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
                        int ks = 0;
                        for (; ks + 16 <= Mp; ks += 16) {
#pragma unroll
                            for (int u = 0; u < 4; u++) dmma884(c[u][0], c[u][1], pa[ks + 4 * u], pb[ks + 4 * u]);
                        }
                        if (ks < Mp) {
#pragma unroll
                            for (int u = 0; u < 4; u++) if (ks + 4 * u < Mp) dmma884(c[u][0], c[u][1], pa[ks + 4 * u], pb[ks + 4 * u]);
                        }
                        const double c0 = ((c[0][0] + c[1][0]) + (c[2][0] + c[3][0])) * sc2, c1 = ((c[0][1] + c[1][1]) + (c[2][1] + c[3][1])) * sc2;
                        const int r = ta * 8 + gq, cc = tb * 8 + 2 * qq;
                        *(double2*)(dst + r * lds + cc) = make_double2(c0, c1);
                        if (ta != tb) { dst[cc * lds + r] = c0; dst[(cc + 1) * lds + r] = c1; }
                        else if ((gq >> 1) == qq) dgd[r] = (gq & 1) ? c1 : c0;
                    }
                    if (it > 0 && wid == nwork - 1) { const double t = warp_diag(dgs); if (lane == 0) trs[it & 1] = t; }   // tr(B_it)
                    asm volatile("bar.sync 1, %0;" ::"r"(nwork * 32) : "memory");
                    bool conv = false;
                    double u;
                    if (it == 0) u = sc * T0;
                    else {
                        const double Tj = trs[it & 1];
                        if (!(Tj > 0.0)) { degenerate = true; break; }
                        conv = Tj > (1.0 - PLS_EIG_DELTA) * u_prev * u_prev;
                        u = sc * Tj;
                    }
                    src = dst; dst = (dst == Sa) ? Sb : Sa;
                    { const double* tswap = dgs; dgs = dgd; dgd = (double*)tswap; }
                    if (g.prof && tid == 0) pacc[5] += 1;
                    if (conv) break;
                    u_prev = u;
                }
                if (tid == 0) { s_flags[0] = degenerate ? 1 : 0; s_flags[1] = (src == Sa) ? 0 : (src == Sb ? 1 : 2); s_flags[2] = (dgs == dgA) ? 0 : 1; }
            }
            __syncthreads();
            if (!deg0) {   // every warp adopts the outcome of the iteration
                degenerate = s_flags[0] != 0;
                src = (s_flags[1] == 0) ? Sa : (s_flags[1] == 1 ? Sb : S0);
                dgs = (s_flags[2] == 0) ? dgA : dgB;
            }
            PROF(6);
            // ---- q (up to scale): one power step with S0 applied to the projector's column with the largest diagonal.
            // The length of q never matters (w is normalised below), so no sqrt / division sits on the chain; the
            // power-of-two sc0 = 2^-exponent(tr S0) only keeps magnitudes O(1).
            if (degenerate) {
                for (int a = tid; a < Mp; a += GT) qv[a] = (a == 0) ? 1.0 : 0.0;
            } else {
                double bv = -1.0; int bi = 0;                  // per warp (same result in every warp): first largest diagonal entry
                for (int a = lane; a < M; a += 32) { const double v = dgs[a]; if (v > bv) { bv = v; bi = a; } }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                }
                matvec(S0, lds, M, M, src + bi, lds, [&](int a, double v) { qv[a] = v * sc0; });
            }
            __syncthreads();
            PROF(1);
        }
        // ---- phase C: w^ = XY q (pls.cpp:408), unnormalised; its squared length is reduced alongside ----------------
        double ww = 0;
        for (int k = tid; k < K; k += GT) {
            double a0 = 0, a1 = 0;
            if (M == 1) a0 = XY[k];                                                             // pls.cpp:403-404
            else {
                int m = 0;
                for (; m + 1 < M; m += 2) { a0 = fma(XY[(size_t)m * ldxy + k], qv[m], a0); a1 = fma(XY[(size_t)(m + 1) * ldxy + k], qv[m + 1], a1); }
                if (m < M) a0 = fma(XY[(size_t)m * ldxy + k], qv[m], a0);
            }
            const double v = a0 + a1;
            wv[k] = v;
            ww = fma(v, v, ww);
        }
        ww = warp_sum(ww);
        if (lane == 0) red[wid] = ww;
        __syncthreads();
        PROF(2);
        // ---- phase D: c_j = P_j^T w^ (pls.cpp:415); r = (w^ - sum_j c_j R_j) / |w^| (pls.cpp:411-416, linear in w) -------------
        matvec(Ps, ldpr, comp, K, wv, 1, [&](int j, double v) { cv[j] = v; });
        double wsq = (lane < GW) ? red[lane] : 0.0;            // one load per warp + shuffles instead of 16 broadcast loads per thread
#pragma unroll
        for (int o = GW / 2; o > 0; o >>= 1) wsq += __shfl_xor_sync(0xffffffffu, wsq, o);
        wsq = __shfl_sync(0xffffffffu, wsq, 0);
        const double inv_wn = rsqrt(wsq);                      // overlaps the matrix-vector product above
        __syncthreads();
        matvec(Rts, ldrt, K, comp, cv, 1, [&](int k, double v) {
            const double wk = wv[k], r = (wk - v) * inv_wn;
            rv[k] = r;
            g.W[(size_t)comp * K + k] = wk * inv_wn;
            g.R[(size_t)comp * K + k] = r;
            Rts[(size_t)k * ldrt + comp] = r;
        });
        __syncthreads();
        PROF(3);
        // ---- phase E: p^ = XX r, tt = r^T XX r (pls.cpp:422-424); q^ = XY^T r (pls.cpp:428) ------------------------------
        double tpart = 0;
        matvec(XX, ldxx, K, K, rv, 1, [&](int b, double v) { pv[b] = v; tpart = fma(v, rv[b], tpart); });
        matvec(XY, ldxy, M, K, rv, 1, [&](int m, double v) { qv[m] = v; });
        tpart = warp_sum(tpart);
        if (lane == 0) red[GW + wid] = tpart;
        __syncthreads();
        double tt = (lane < GW) ? red[GW + lane] : 0.0;
#pragma unroll
        for (int o = GW / 2; o > 0; o >>= 1) tt += __shfl_xor_sync(0xffffffffu, tt, o);
        tt = __shfl_sync(0xffffffffu, tt, 0);
        const double inv_tt = 1.0 / tt;
        for (int k = tid; k < K; k += GT) {                                                    // pls.cpp:427
            const double p = pv[k] * inv_tt;
            g.P[(size_t)comp * K + k] = p;
            if (g.pr_smem) Ps[(size_t)comp * ldpr + k] = p;
        }
        for (int m = tid; m < M; m += GT) g.Q[(size_t)comp * M + m] = qv[m] * inv_tt;         // pls.cpp:428
        for (int i = tid; i < K * M; i += GT) {                                                // pls.cpp:429
            const int m = i / K, k = i - m * K;
            XY[(size_t)m * ldxy + k] -= ((pv[k] * inv_tt) * (qv[m] * inv_tt)) * tt;
        }
        __syncthreads();
        PROF(4);
    }
    if (g.prof && tid == 0) for (int i = 0; i < 8; i++) g.prof[i] = pacc[i];
#undef PROF
}

}  // namespace

size_t pls_gram_ws_bytes(const abcb200_ctx* ctx, int64_t n, int K, int M) {
    return 2 * align_up((size_t)K * M * 8, 256) + align_up((size_t)K * K * 8, 256) + 512 + align_up((size_t)K * K * 8, 256) + gram_ws_bytes(ctx, n, K, M) + 1024 + pls_defl_ws_bytes(K, K) + pls_wide_ws_bytes(K, M, K);
}

// Fits f.A components from X (n x K), Y (n x M): two Gram products (DMMA) + the persistent component-loop CTA.
// Fills W, P, R, Q; T is left to the caller (T = X R, pls.cpp:418 — launch_xb).
int pls_fit_gram_dev(abcb200_ctx* ctx, const double* X, int64_t ldx, const double* Y, int64_t ldy, const PlsFactors& f) {
    const int K = f.K, M = f.M;
    const int64_t n = f.n;
    double* XY = ws_new<double>(ctx, (size_t)K * M);
    double* XX = ws_new<double>(ctx, (size_t)K * K);
    if (!XY || !XX) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in pls_fit_gram");
    ABC_TRY(launch_gram(ctx, X, ldx, K, Y, ldy, M, n, XX, XY));   // pls.cpp:396, :398 (kernel timer 1 inside)
    return pls_components_dev(ctx, XX, XY, f);
}

size_t pls_components_ws_bytes(int K, int M, int A) {
    return align_up((size_t)K * M * 8, 256) + align_up((size_t)K * A * 8, 256) + 512 + pls_defl_ws_bytes(K, A) + pls_wide_ws_bytes(K, M, A) + 1024;
}

// The component loop alone, from XX = X^T X (K x K) and XY = X^T Y (K x M, ld K): fills W, P, R, Q (pls.cpp:400-435).
int pls_components_dev(abcb200_ctx* ctx, const double* XX, const double* XY, const PlsFactors& f) {
    const int K = f.K, M = f.M, A = f.A;
    if (M > 128) ABC_FAIL(ctx, ABCB200_EINVAL, "pls_fit: M=%d responses exceed the on-chip eigen-solver limit (128)", M);
    double* XYg = ws_new<double>(ctx, (size_t)K * M);
    double* Rt = ws_new<double>(ctx, (size_t)K * A);
    long long* prof = ws_new<long long>(ctx, 8);
    if (!XYg || !Rt || !prof) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in pls_components");
    static const bool want_prof = getenv("ABCB200_PLS_PROF") != nullptr;     // debug: per-phase clock totals on stderr
    static const bool literal = getenv("ABCB200_PLS_LITERAL") != nullptr;    // force the R/P-recurrence loop below
    if (!literal && pls_defl_fits(ctx, K, M)) {                              // deflated-Gram loop, everything on chip (pls_defl.cu)
        if (want_prof) CUDA_TRY(ctx, cudaMemsetAsync(prof, 0, 8 * sizeof(long long), ctx->stream));
        ctx->stat_pls_loop = 1;
        ABC_TRY(pls_defl_dev(ctx, XX, XY, f, want_prof ? prof : nullptr));
        if (want_prof) {
            long long h[8];
            CUDA_TRY(ctx, cudaMemcpyAsync(h, prof, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            fprintf(stderr, "[pls_defl K=%d M=%d A=%d] cycles per component: S0 %.0f | squarings %.0f (%.1f iterations; warp 0: tiles %.0f, barrier wait %.0f) | w %.0f | H update, p, tt, q %.0f | deflate XY %.0f\n", K, M, A,
                    (double)h[0] / A, (double)h[6] / A, (double)h[5] / A, (double)h[1] / A, (double)h[7] / A, (double)h[2] / A, (double)h[3] / A, (double)h[4] / A);
        }
        return ABCB200_OK;
    }
    if (!literal && pls_wide_fits(ctx, K, M)) {                              // wide predictor sets: H and XY in L2, whole-GPU launches (pls_wide.cu)
        ctx->stat_pls_loop = 3;
        return pls_wide_dev(ctx, XX, XY, f);
    }
    GramArgs g;
    g.XX = XX; g.XY0 = XY; g.XYg = XYg; g.W = f.W; g.P = f.P; g.R = f.R; g.Q = f.Q; g.Rt = Rt; g.K = K; g.M = M; g.A = A;
    int ldk = K;
    while (ldk % 16 != 4) ldk++;
    g.ldk = ldk;
    int lda = A;
    while (lda % 16 != 4) lda++;
    g.lda = lda;
    g.prof = want_prof ? prof : nullptr;
    if (want_prof) CUDA_TRY(ctx, cudaMemsetAsync(prof, 0, 8 * sizeof(long long), ctx->stream));
    const size_t Mp = (size_t)(M + 7) / 8 * 8;
    const size_t fixed = sizeof(double) * (3 * Mp * (Mp + 4) + 3 * Mp + 3 * (size_t)K + A + 34) + 256;
    const size_t budget = (size_t)ctx->smem_optin;
    if (fixed > budget) ABC_FAIL(ctx, ABCB200_EINVAL, "pls_fit: K=%d M=%d A=%d need %zu B of shared memory (> %zu)", K, M, A, fixed, budget);
    size_t used = fixed;
    const size_t xy_b = (size_t)M * ldk * 8, xx_b = (size_t)K * ldk * 8, pr_b = ((size_t)A * ldk + (size_t)K * lda) * 8;
    g.xy_smem = (used + xy_b <= budget) ? 1 : 0; if (g.xy_smem) used += xy_b;
    g.xx_smem = (used + xx_b <= budget) ? 1 : 0; if (g.xx_smem) used += xx_b;
    g.pr_smem = (used + pr_b <= budget) ? 1 : 0; if (g.pr_smem) used += pr_b;
    CUDA_TRY(ctx, cudaFuncSetAttribute(pls_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)used));
    ctx->stat_pls_loop = 2;
    kernel_begin(ctx, 0);
    LAUNCH(ctx, pls_gram_kernel, 1, GT, used, g);
    kernel_end(ctx, 0);
    if (want_prof) {
        long long h[8];
        CUDA_TRY(ctx, cudaMemcpyAsync(h, prof, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        fprintf(stderr, "[pls_gram K=%d M=%d A=%d smem xy/xx/pr=%d/%d/%d] cycles per component: S0 %.0f | squarings %.0f (%.1f iterations) | extract q %.0f | w %.0f | c,r %.0f | p,q,deflate %.0f\n", K, M, A,
                g.xy_smem, g.xx_smem, g.pr_smem, (double)h[0] / A, (double)h[6] / A, (double)h[5] / A, (double)h[1] / A, (double)h[2] / A, (double)h[3] / A, (double)h[4] / A);
    }
    return ABCB200_OK;
}
