// pls_defl.cu — the component loop of kernel PLS with EVERYTHING on chip, one persistent CTA.
//
// Reference: PLS::Model::plsr, lib/PLS/src/pls.cpp:400-435. Per component the reference needs the dominant
// eigenvector q of XY^T XY (:406), w = XY q normalised (:408-411), r = w - sum_j (p_j^T w) r_j (:412-416), t = X r,
// tt = t^T t, p = X^T t / tt, q = Y^T t / tt (:418-428) and the deflation XY -= tt p q^T (:429).
//
// pls_gram.cu evaluates those formulas literally from XX = X^T X and XY; its r-recurrence and p = XX r stream
// K x a and K x K operands that do not fit in shared memory from L2 twice per component (25 us per component at
// K = 150). Here the same quantities are obtained from the DEFLATED Gram matrix H_a = X_a^T X_a (X_a = X with the
// first a score directions projected out; H_0 = XX):
//       t = X r = X_a w        =>  tt = w^T H_a w,  p = H_a w / tt,  q = XY_a^T w / tt        (XY_a = X_a^T Y, :429)
//       X_{a+1} = X_a (I - w p^T)  =>  H_{a+1} = H_a - tt p p^T                               (rank one, symmetric)
// so a component never touches R or P of the earlier components and H (packed upper triangle) + XY stay in shared
// memory for the whole fit. R follows afterwards from the reference's own recurrence r_a = w_a - sum_j (p_j^T w_a) r_j
// (:412-416) with U = P^T W formed once (pls_u_kernel, pls_r_kernel: row-parallel, blocked by 8 columns).
// The normalisation of w is not on the critical path either: with w^ = XY q (any length), p^ = H w^, q^ = XY^T w^,
// tt^ = w^^T p^ the deflations are p^ p^^T / tt^ and p^ q^^T / tt^ exactly, and W = w^/|w^|, P = p^ |w^|/tt^,
// Q = q^ |w^|/tt^ are written by the warps that idle during the next component's eigen-iteration.
//
//   phase A  S0 = XY^T XY (M x M)                       DMMA 8x8 tiles of the upper triangle, mirrored
//   phase B  projector onto the dominant eigenvector    B <- (s B)^2 by DMMA, s a power of two; the trace (scale,
//                                                       convergence) and the arg-max of the diagonal (which column
//                                                       to read q from) are produced by one otherwise idle warp
//                                                       while the next squaring runs
//   phase C  w^ = XY q
//   phase E  H <- H - (pending rank-one term), p^ = H w^ (each stored element used for its row and its column),
//            tt^, q^ = XY^T w^
//   phase F  p^ from the partial sums, XY -= p^ q^^T / tt^
//
// Tried and dropped (round 2, measured at C3): (1) letting the warps that idle during the eigen-iteration apply the pending rank-one
// term to H so that phase E only reads it: their shared-memory traffic slows the latency-critical squaring warps, the loop went from
// 2.04 to 2.41 ms; (2) ending the squarings early and finishing with single-warp products v <- B v (kernels.cuh, PLS_EIG_DELTA).
// Differences from the literal formulas are rounding-level (1e-13 on coefficients, measured against the oracle in
// tests/test_gpu_parity.py::test_pls_model). Shapes whose H does not fit (K > ~170) use pls_gram.cu.
#include <stdlib.h>

#include "kernels.cuh"

namespace {

constexpr int DT = 512;             // threads of the persistent CTA
constexpr int DW = DT / 32;

struct DeflArgs {
    const double* XX;   // K x K (ld K), symmetric
    const double* XY0;  // K x M (ld K)
    double *W, *P, *Q;  // K x A, K x A, M x A
    long long* prof;    // optional per-phase clock totals (debug), 8 entries
    int K, M, A, ldk;
    // leave-one-out mode (Model::cv_LOO, pls.cpp:469-491): one refit per held-out row, nothing but the residuals leaves the SM
    const double* X;    // N x K (ldx): the rows the model holds (_X)
    const double* Y;    // N x M (ldy)
    double* cube;       // M x (N x A): Ev[y](row, comp) at cube[(y * A + comp) * N + row] (pls.cpp:474, 479-480)
    long long N, ldx, ldy;
    // chunked fit (pipelined ranking, api.cu): this launch runs components [c0, c1); everything the next launch needs travels
    // through `state` (H with its pending rank-one term = p^ and 1 / tt^ of the last component, the deflated XY)
    int c0, c1;
    double* state;      // (K (K + 1) / 2 + M ldk + Kp + 4) doubles, or null when c0 == 0 and c1 == A
};

// Shared-memory accesses through an opaque 32-bit shared address. Inside the eigen-iteration the compiler otherwise
// re-derives the shared window base (S2R SR_CgaCtaId + LEA) in front of every access to a static or carved-out array, on
// the critical path of a loop whose whole body is a few hundred cycles.
__device__ __forceinline__ unsigned smem_opaque(const void* p) {
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("" : "+r"(a));
    return a;
}
__device__ __forceinline__ double lds_f64(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ int lds_s32(unsigned a) { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_f64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_s32(unsigned a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

__device__ __forceinline__ double pow2_inv(double x) {     // 2^-exponent(x): x * result in [1, 2)
    const int ex = ((__double2hiint(x) >> 20) & 0x7ff) - 1023;
    return __hiloint2double((1023 - ex) << 20, 0);
}

// KS = ceil(K / 32): column slots per lane (and KS * 2 rows per warp)
// LOO = false: one CTA, one fit, factors written to W / P / Q.
// LOO = true : grid-stride over held-out rows. The training Gram matrices of row i are rank-one down-dates of the full
//   ones (XX - x_i x_i^T, XY - x_i y_i^T: a sum over rows does not care that the reference keeps the training rows in a
//   permuted order, pls.cpp:484-486), the refit runs on chip as usual, and instead of the factors the CTA keeps the held-out
//   row deflated alongside: with s_a = (x^(a) . w^_a) / tt^_a,  x^(a+1) = x^(a) - s_a p^_a  and  e_(a+1) = e_a - s_a q^_a
//   (x^(0) = x_i, e_0 = y_i), e_(a+1) is exactly y_i - x_i R[:, :a+1] Q[:, :a+1]^T (pls.cpp:449-455), O(K + M) per component.
template <int KS, bool LOO>
__global__ void __launch_bounds__(DT, 1) pls_defl_kernel(DeflArgs g) {
    constexpr int Kp = 32 * KS;
    constexpr int RMAX = 2 * KS;
    extern __shared__ __align__(16) double sm[];
    const int K = g.K, M = g.M, A = g.A, ldk = g.ldk;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int gq = lane >> 2, qq = lane & 3;
    const int Mp = (M + 7) / 8 * 8, lds = Mp + 4, ntile = Mp / 8;
    const int ssz = Mp * lds;
    const int rsz = max(3 * ssz, 17 * Kp);      // S0 | Sa | Sb, re-used for the 16 row partials per row of phase E (row stride 17: see rowp)
    double* S0 = sm;
    double* Sa = S0 + ssz;
    double* Sb = Sa + ssz;
    double* rowp = sm;                // [Kp][17] during phases E/F: 16 partials per row, stride 17 doubles so that phase F's
                                      // thread-per-row reads are bank-conflict free (stride 16 put every lane on the same bank)
    double* dgA = sm + rsz;           // Mp: compact diagonals of the eigen iterates (ping-pong)
    double* dgB = dgA + Mp;           // Mp
    double* trs = dgB + Mp;           // 4: traces of the eigen iterates (ping-pong) | arg-max of the diagonal
    double* qh = trs + 4;             // Mp: q^
    double* wv = qh + Mp;             // Kp: w^ (padding stays 0)
    double* ph = wv + Kp;             // Kp: p^ of the last finished component = pending rank-one term of H
    double* colp = ph + Kp;           // DW * Kp: per-warp column partial sums
    double* ttp = colp + DW * Kp;     // DW
    double* wwp = ttp + DW;           // DW
    double* scal = wwp + DW;          // 4: [0] 1 / tt^ of the last finished component
    double* xc = scal + 4;            // LOO: Kp, the held-out row deflated by the finished components
    double* ec = xc + (LOO ? Kp : 0); // LOO: Mp, its residual
    double* XY = ec + (LOO ? Mp : 0); // M x ldk
    double* H = XY + (size_t)M * ldk; // packed upper triangle, row i at i*K - i(i-1)/2, entries (i, i..K-1)
    __shared__ int s_flags[4];
    __shared__ int s_amax[2];
    __shared__ unsigned char pair_ta[136], pair_tb[136];      // upper-triangle tile pairs, ntile <= 16
    const int npair = ntile * (ntile + 1) / 2;
    if (tid == 0) { int p = 0; for (int ta = 0; ta < ntile; ta++) for (int tb = ta; tb < ntile; tb++) { pair_ta[p] = (unsigned char)ta; pair_tb[p] = (unsigned char)tb; p++; } }
    long long tprev = clock64(), pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // per-phase clock totals of thread 0, kept in registers
#define PROF(slot) do { if (!LOO && g.prof && tid == 0) { const long long tn = clock64(); pacc[slot] += tn - tprev; tprev = tn; } } while (0)
    for (long long row = LOO ? (long long)blockIdx.x : 0; row < (LOO ? g.N : 1); row += gridDim.x) {
    for (int i = tid; i < (int)(H - sm); i += DT) sm[i] = 0.0;
    __syncthreads();
    if (LOO) {
        for (int k = tid; k < K; k += DT) xc[k] = g.X[(size_t)k * g.ldx + row];
        for (int m = tid; m < M; m += DT) ec[m] = g.Y[(size_t)m * g.ldy + row];
        __syncthreads();
    }
    const int c_first = LOO ? 0 : g.c0, c_last = LOO ? A : g.c1;
    const int nH = K * (K + 1) / 2;
    if (!LOO && c_first > 0) {        // resume: what the previous launch left in `state`
        const double* st = g.state;
        for (int i = tid; i < nH; i += DT) H[i] = st[i];
        for (int i = tid; i < M * ldk; i += DT) XY[i] = st[nH + i];
        for (int i = tid; i < Kp; i += DT) ph[i] = st[nH + M * ldk + i];
        if (tid == 0) scal[0] = st[nH + M * ldk + Kp];
    } else {
    for (int i = tid; i < K * M; i += DT) {
        const int m = i / K, k = i - m * K;
        XY[(size_t)m * ldk + k] = LOO ? fma(-xc[k], ec[m], g.XY0[i]) : g.XY0[i];
    }
    for (int i = tid; i < K * K; i += DT) {
        const int r = i / K, c = i - r * K;
        if (c >= r) H[r * K - r * (r - 1) / 2 + (c - r)] = LOO ? fma(-xc[r], xc[c], g.XX[(size_t)c * K + r]) : g.XX[(size_t)c * K + r];
    }
    }
    __syncthreads();

    // rows of H this warp owns in phase E: serpentine over blocks of 32 (balanced lengths)
    auto row_of = [&](int t) { return 32 * (t >> 1) + ((t & 1) ? 31 - wid - 16 * 0 : wid) + ((t & 1) ? 0 : 0); };
    const int nwork = min(DW, npair + 1);       // warps that take part in the eigen-iteration (tile pairs + the trace warp)

    // W, P, Q of a finished component from w^, p^, q^, |w^|^2 partials and 1 / tt^ (pls.cpp:411, 427, 428)
    auto emit = [&](int comp, int t0, int nt) {
        if (LOO) {          // one warp (t0 < 32 of it) carries the held-out row through the finished component
            if (t0 >= 32) return;
            double d = 0.0;
            for (int k = t0; k < K; k += 32) d = fma(xc[k], wv[k], d);
            d = warp_sum(d);
            const double sc = d * scal[0];
            for (int k = t0; k < K; k += 32) xc[k] = fma(-sc, ph[k], xc[k]);
            for (int m = t0; m < M; m += 32) {
                const double e = fma(-sc, qh[m], ec[m]);
                ec[m] = e;
                g.cube[((size_t)m * A + comp) * (size_t)g.N + (size_t)row] = e;
            }
            return;
        }
        double ww = 0.0;
        for (int w = 0; w < DW; w++) ww += wwp[w];
        const double n = sqrt(ww), f = n * scal[0];
        for (int k = t0; k < K; k += nt) {
            g.W[(size_t)comp * K + k] = wv[k] / n;
            g.P[(size_t)comp * K + k] = ph[k] * f;
        }
        for (int m = t0; m < M; m += nt) g.Q[(size_t)comp * M + m] = qh[m] * f;
    };

    for (int comp = c_first; comp < c_last; comp++) {
        int bi = 0;
        bool degenerate = false;
        const double* src = S0;
        if (M != 1) {
            // ---- phase A: S0 = XY^T XY (pls.cpp:406), upper-triangle tiles mirrored (bitwise symmetric) ---------
            for (int pidx = wid; pidx < npair; pidx += DW) {
                const int ta = pair_ta[pidx], tb = pair_tb[pidx];
                const int ca = ta * 8 + gq, cb = tb * 8 + gq;
                const bool va = ca < M, vb = cb < M;
                const double* pa = XY + (size_t)min(ca, M - 1) * ldk;
                const double* pb = XY + (size_t)min(cb, M - 1) * ldk;
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
            // ---- phase B: projector onto the dominant eigenvector by trace-normalised repeated squaring -------
            // B_{j+1} = (s_j B_j)^2 with s_j a power of two (exact scaling). With u_j = s_j tr(B_j):
            // tr(B_{j+1}) / u_j^2 = sum l_i^2 / (sum l_i)^2 -> 1 exactly when B_j has rank one, so
            // "tr(B_{j+1}) > (1 - d) u_j^2" says B_j had l2/l1 < d/2; noticed one squaring late, the iterate used is B_{j+2}:
            // ratios (d/2)^4 (PLS_EIG_DELTA, kernels.cuh).
            // Trace and arg-max of the diagonal of iterate j are produced by warp nwork-1 while squaring j runs; the
            // scale of step j comes from the bound tr(B_j) <= u_{j-1}^2 (within a factor M of the truth, re-centred
            // every step) and convergence is noticed one squaring late, which costs nothing in accuracy.
            auto warp_diag = [&](const double* dg) {
                double t = 0;
                for (int a = lane; a < Mp; a += 32) t += dg[a];
                return warp_sum(t);
            };
            double* dst = Sa;
            const double* dgs = dgA;
            double* dgd = dgB;
            const double T0 = warp_diag(dgs);                 // same bits in every warp
            const bool deg0 = !(T0 > 0.0) || !(T0 < 1e300);   // zero, NaN or inf matrix (uniform over the CTA)
            degenerate = deg0;
            if (!degenerate && wid < nwork) {
                double u_prev = 0.0;
                const int ta_first = pair_ta[min(wid, npair - 1)], tb_first = pair_tb[min(wid, npair - 1)];   // this warp's first (usually only) tile pair
                const unsigned trs_a = smem_opaque(trs), amax_a = smem_opaque(s_amax);
                for (int it = 0; it < 80; it++) {
                    const double sc = (it == 0) ? pow2_inv(T0) : pow2_inv(u_prev * u_prev);
                    const double sc2 = sc * sc;
                    long long tq0 = 0;
                    if (!LOO && g.prof && tid == 0) tq0 = clock64();
                    for (int pidx = wid; pidx < npair; pidx += DW) {
                        const int ta = (pidx == wid) ? ta_first : pair_ta[pidx], tb = (pidx == wid) ? tb_first : pair_tb[pidx];
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
                        if (lane == 0) { sts_f64(trs_a + 8 * (it & 1), t); sts_s32(amax_a + 4 * (it & 1), bj); }
                    }
                    if (!LOO && g.prof && tid == 0) { const long long tq = clock64(); pacc[1] += tq - tq0; tq0 = tq; }
                    asm volatile("bar.sync 1, %0;" ::"r"(nwork * 32) : "memory");
                    if (!LOO && g.prof && tid == 0) { const long long tq = clock64(); pacc[7] += tq - tq0; }
                    bool conv = false;
                    double u;
                    if (it == 0) u = sc * T0;
                    else {
                        const double Tj = lds_f64(trs_a + 8 * (it & 1));
                        bi = lds_s32(amax_a + 4 * (it & 1));
                        if (!(Tj > 0.0)) { degenerate = true; break; }
                        conv = Tj > (1.0 - PLS_EIG_DELTA) * u_prev * u_prev;
                        u = sc * Tj;
                    }
                    src = dst; dst = (dst == Sa) ? Sb : Sa;
                    { const double* tswap = dgs; dgs = dgd; dgd = (double*)tswap; }
                    if (!LOO && g.prof && tid == 0) pacc[5] += 1;
                    if (conv) break;
                    u_prev = u;
                }
                if (tid == 0) { s_flags[0] = degenerate ? 1 : 0; s_flags[1] = (src == Sa) ? 0 : (src == Sb ? 1 : 2); s_flags[2] = bi; }
            } else if (wid >= nwork && comp > c_first) {
                emit(comp - 1, (wid - nwork) * 32 + lane, (DW - nwork) * 32);      // idle warps: outputs of the previous component
            }
            if (nwork == DW && comp > c_first && !deg0) { /* no idle warp: outputs are written after the barrier below */ }
            __syncthreads();
            if (!deg0) {   // every warp adopts the outcome of the iteration
                degenerate = s_flags[0] != 0;
                src = (s_flags[1] == 0) ? Sa : (s_flags[1] == 1 ? Sb : S0);
                bi = s_flags[2];
            }
            if ((nwork == DW || deg0) && comp > c_first) emit(comp - 1, tid, DT);      // (deg0: nobody ran the idle-warp branch's twin)
            PROF(6);
        } else if (comp > c_first) {
            emit(comp - 1, tid, DT);
        }
        // ---- phase C: w^ = XY q (pls.cpp:408), unnormalised; q = column bi of the projector ------------------------
        // (emit() above read wv / ph / qh / wwp of the previous component: order the overwrite after it)
        if (M == 1 || nwork == DW || degenerate) __syncthreads();
        {
            double ww = 0.0;
            for (int k = tid; k < K; k += DT) {
                double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
                if (M == 1) a0 = XY[k];                                                             // pls.cpp:403-404
                else if (degenerate) a0 = XY[k];                                                    // q = e_0
                else {
                    const double* qc = src + bi * lds;
                    int m = 0;
                    for (; m + 3 < M; m += 4) {
                        a0 = fma(XY[(size_t)m * ldk + k], qc[m], a0); a1 = fma(XY[(size_t)(m + 1) * ldk + k], qc[m + 1], a1);
                        a2 = fma(XY[(size_t)(m + 2) * ldk + k], qc[m + 2], a2); a3 = fma(XY[(size_t)(m + 3) * ldk + k], qc[m + 3], a3);
                    }
                    for (; m < M; m++) a0 = fma(XY[(size_t)m * ldk + k], qc[m], a0);
                }
                const double v = (a0 + a1) + (a2 + a3);
                wv[k] = v;
                ww = fma(v, v, ww);
            }
            ww = warp_sum(ww);
            if (lane == 0) wwp[wid] = ww;
        }
        __syncthreads();
        PROF(2);
        // ---- phase E: apply the pending rank-one term to H, p^ = H w^ from the packed triangle, tt^, q^ = XY^T w^ ----
        {
            double wj[KS], pj[KS], colacc[KS];
#pragma unroll
            for (int c = 0; c < KS; c++) { wj[c] = wv[32 * c + lane]; pj[c] = ph[32 * c + lane]; colacc[c] = 0.0; }
            const double pscale = scal[0];
            double tl = 0.0;
#pragma unroll
            for (int t = 0; t < RMAX; t++) {
                const int i = 32 * (t >> 1) + ((t & 1) ? 31 - wid : wid);     // serpentine: balanced row lengths per warp
                if (i < K) {                                                  // warp-uniform
                    const double wi = wv[i], pis = ph[i] * pscale;
                    double* hrow = H + (i * K - i * (i - 1) / 2) - i;         // hrow[j] = H(i, j), j >= i
                    double racc = 0.0;
#pragma unroll
                    for (int c = t >> 1; c < KS; c++) {
                        const int j = 32 * c + lane;
                        if (j >= i && j < K) {
                            const double x = fma(-pis, pj[c], hrow[j]);
                            hrow[j] = x;
                            racc = fma(x, wj[c], racc);
                            if (j > i) colacc[c] = fma(x, wi, colacc[c]);
                        }
                    }
                    racc += __shfl_xor_sync(0xffffffffu, racc, 16);
                    if (lane < 16) rowp[i * 17 + lane] = racc;
                    // sum_i w_i row_i: lanes 0..15 hold the 16 partials of row i
                    tl = fma(lane < 16 ? racc : 0.0, wi, tl);
                }
            }
#pragma unroll
            for (int c = 0; c < KS; c++) { colp[wid * Kp + 32 * c + lane] = colacc[c]; tl = fma(colacc[c], wj[c], tl); }
            tl = warp_sum(tl);
            if (lane == 0) ttp[wid] = tl;
            for (int m0 = wid; m0 < M; m0 += 2 * DW) {      // q^: two responses per trip (independent reductions)
                const int m1 = m0 + DW;
                const double* x0 = XY + (size_t)m0 * ldk;
                const double* x1 = XY + (size_t)min(m1, M - 1) * ldk;
                double q0 = 0.0, q1 = 0.0;
#pragma unroll
                for (int c = 0; c < KS; c++) {
                    const int k = 32 * c + lane;
                    if (k < K) { q0 = fma(x0[k], wj[c], q0); q1 = fma(x1[k], wj[c], q1); }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { q0 += __shfl_xor_sync(0xffffffffu, q0, o); q1 += __shfl_xor_sync(0xffffffffu, q1, o); }
                if (lane == 0) { qh[m0] = q0; if (m1 < M) qh[m1] = q1; }
            }
        }
        __syncthreads();
        PROF(3);
        // ---- phase F: p^ from the partial sums; XY -= p^ q^^T / tt^ (pls.cpp:429) ---------------------------------------
        {
            double tt = 0.0;
#pragma unroll
            for (int w = 0; w < DW; w++) tt += ttp[w];
            const double inv_tt = 1.0 / tt;
            const int i = tid % Kp, part = tid / Kp, nparts = DT / Kp;
            if (i < K && part < nparts) {
                double r0 = 0, r1 = 0, c0 = 0, c1 = 0;
#pragma unroll
                for (int u = 0; u < 16; u += 2) { r0 += rowp[i * 17 + u]; r1 += rowp[i * 17 + u + 1]; }
#pragma unroll
                for (int w = 0; w < DW; w += 2) { c0 += colp[w * Kp + i]; c1 += colp[(w + 1) * Kp + i]; }
                const double p = (r0 + r1) + (c0 + c1);
                const double ps = p * inv_tt;
                for (int m = part; m < M; m += nparts) XY[(size_t)m * ldk + i] = fma(-ps, qh[m], XY[(size_t)m * ldk + i]);
                if (part == 0) ph[i] = p;
            }
            if (tid == 0) scal[0] = inv_tt;
        }
        __syncthreads();
        PROF(4);
    }
    emit(c_last - 1, tid, DT);
    if (LOO) __syncthreads();       // the next held-out row re-initialises everything emit() just read
    if (!LOO && c_last < A) {       // hand over to the launch that continues with component c_last
        double* st = g.state;
        for (int i = tid; i < nH; i += DT) st[i] = H[i];
        for (int i = tid; i < M * ldk; i += DT) st[nH + i] = XY[i];
        for (int i = tid; i < Kp; i += DT) st[nH + M * ldk + i] = ph[i];
        if (tid == 0) st[nH + M * ldk + Kp] = scal[0];
    }
    }
    if (!LOO && g.prof && tid == 0) for (int i = 0; i < 8; i++) g.prof[i] += pacc[i];
#undef PROF
}

// U[j, a] = p_j^T w_a for j < a (pls.cpp:415), column a at U + a * A. One CTA per a, one warp per four j.
__global__ void __launch_bounds__(256) pls_u_kernel(const double* __restrict__ P, const double* __restrict__ W, int K, int A, int a_begin, double* __restrict__ U) {
    extern __shared__ double wa[];
    const int a = a_begin + blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int k = tid; k < K; k += 256) wa[k] = W[(size_t)a * K + k];
    __syncthreads();
    for (int j0 = wid * 4; j0 < a; j0 += 32) {
        double s[4] = {0, 0, 0, 0};
        for (int k = lane; k < K; k += 32) {
            const double w = wa[k];
#pragma unroll
            for (int u = 0; u < 4; u++) s[u] = fma(P[(size_t)min(j0 + u, a - 1) * K + k], w, s[u]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int u = 0; u < 4; u++) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
        }
        if (lane < 4 && j0 + lane < a) U[(size_t)a * A + j0 + lane] = (lane == 0) ? s[0] : (lane == 1) ? s[1] : (lane == 2) ? s[2] : s[3];
    }
}

// R[k, a] = W[k, a] - sum_{j < a} R[k, j] U[j, a] (pls.cpp:412-416): rows are independent, one warp per row, eight
// columns per trip (the long dot products of a trip are independent; the 8 x 8 triangle inside it is solved serially).
// Columns [a_begin, a_end) only (a_begin a multiple of 8); the earlier columns of the row are read back from R.
__global__ void __launch_bounds__(256) pls_r_kernel(const double* __restrict__ W, const double* __restrict__ U, int K, int A, int a_begin, int a_end,
                                                    double* __restrict__ R) {
    extern __shared__ double rrow_all[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int k = blockIdx.x * 8 + wid;
    if (k >= K) return;
    double* rrow = rrow_all + (size_t)wid * A;
    for (int j = lane; j < a_begin; j += 32) rrow[j] = R[(size_t)j * K + k];
    __syncwarp();
    for (int a0 = a_begin; a0 < a_end; a0 += 8) {
        double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int j = lane; j < a0; j += 32) {
            const double rj = rrow[j];
#pragma unroll
            for (int b = 0; b < 8; b++) acc[b] = fma(rj, U[(size_t)min(a0 + b, A - 1) * A + j], acc[b]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int b = 0; b < 8; b++) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], o);
        }
        double r[8];
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const int a = a0 + b;
            if (a < a_end) {
                double v = W[(size_t)a * K + k] - acc[b];
#pragma unroll
                for (int b2 = 0; b2 < b; b2++) v = fma(-r[b2], U[(size_t)a * A + a0 + b2], v);
                r[b] = v;
            } else r[b] = 0.0;
        }
        __syncwarp();
        if (lane < 8 && a0 + lane < a_end) {
            double v = r[0];
#pragma unroll
            for (int b = 1; b < 8; b++) if (lane == b) v = r[b];
            rrow[a0 + lane] = v;
            R[(size_t)(a0 + lane) * K + k] = v;
        }
        __syncwarp();
    }
}

size_t defl_smem_doubles(int K, int M, bool loo = false) {
    const size_t Mp = (size_t)(M + 7) / 8 * 8, Kp = (size_t)(K + 31) / 32 * 32;
    int ldk = K;
    while (ldk % 16 != 4 && ldk % 16 != 12) ldk++;
    const size_t rsz = (3 * Mp * (Mp + 4) > 17 * Kp) ? 3 * Mp * (Mp + 4) : 17 * Kp;
    return rsz + 3 * Mp + 4 + 2 * Kp + DW * Kp + 2 * DW + 4 + (loo ? Kp + Mp : 0) + (size_t)M * ldk + (size_t)K * (K + 1) / 2;
}

}  // namespace

// R from W and P by the reference's recurrence r_a = w_a - sum_j (p_j^T w_a) r_j (pls.cpp:412-416); U: A x A scratch
int pls_ur_dev(abcb200_ctx* ctx, const PlsFactors& f, double* U) { return pls_ur_block_dev(ctx, f, U, 0, f.A); }

// the same for columns [a_begin, a_end) once the earlier columns of R exist (a_begin a multiple of 8)
int pls_ur_block_dev(abcb200_ctx* ctx, const PlsFactors& f, double* U, int a_begin, int a_end) {
    const int K = f.K, A = f.A;
    if (a_end > a_begin && a_end > 1) LAUNCH(ctx, pls_u_kernel, a_end - a_begin, 256, (size_t)K * 8, f.P, f.W, K, A, a_begin, U);
    LAUNCH(ctx, pls_r_kernel, (K + 7) / 8, 256, (size_t)8 * A * 8, f.W, U, K, A, a_begin, a_end, f.R);
    return ABCB200_OK;
}

// true when the all-on-chip component loop can take this shape
bool pls_defl_fits(const abcb200_ctx* ctx, int K, int M) {
    return K <= 192 && M <= 128 && defl_smem_doubles(K, M) * 8 + 512 <= (size_t)ctx->smem_optin;
}

size_t pls_defl_ws_bytes(int K, int A) { return align_up((size_t)A * A * 8, 256) + 512; }

// doubles the chunked fit hands from one launch to the next (DeflArgs::state)
size_t pls_defl_state_doubles(int K, int M) {
    int ldk = K;
    while (ldk % 16 != 4 && ldk % 16 != 12) ldk++;
    return (size_t)K * (K + 1) / 2 + (size_t)M * ldk + (size_t)(K + 31) / 32 * 32 + 4;
}

// Components [c0, c1) of the fit (c0 a multiple of 8): W, P, Q columns of the range are written; R is the caller's business
// (pls_ur_block_dev). `state` carries the loop from one launch to the next; runs on ctx->stream.
int pls_defl_chunk_dev(abcb200_ctx* ctx, const double* XX, const double* XY, const PlsFactors& f, int c0, int c1, double* state, long long* prof) {
    const int K = f.K, M = f.M, A = f.A;
    DeflArgs g;
    g.XX = XX; g.XY0 = XY; g.W = f.W; g.P = f.P; g.Q = f.Q; g.prof = prof; g.K = K; g.M = M; g.A = A;
    g.X = g.Y = nullptr; g.cube = nullptr; g.N = g.ldx = g.ldy = 0;
    g.c0 = c0; g.c1 = c1; g.state = state;
    int ldk = K;
    while (ldk % 16 != 4 && ldk % 16 != 12) ldk++;
    g.ldk = ldk;
    const size_t smem = defl_smem_doubles(K, M) * 8;
    const int KS = (K + 31) / 32;
#define DEFL_CASE(KS_)                                                                                                             \
    case KS_: {                                                                                                                    \
        auto kfn = pls_defl_kernel<KS_, false>;                                                                                    \
        CUDA_TRY(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                          \
        LAUNCH(ctx, kfn, 1, DT, smem, g);                                                                                          \
        break;                                                                                                                     \
    }
    switch (KS) {
        DEFL_CASE(1) DEFL_CASE(2) DEFL_CASE(3) DEFL_CASE(4) DEFL_CASE(5) DEFL_CASE(6)
        default: ABC_FAIL(ctx, ABCB200_EINVAL, "pls_defl: K=%d too large", K);
    }
#undef DEFL_CASE
    return ABCB200_OK;
}

// Component loop from XX (K x K) and XY (K x M): fills W, P, Q and R (ld K / M as in PlsFactors).
int pls_defl_dev(abcb200_ctx* ctx, const double* XX, const double* XY, const PlsFactors& f, long long* prof) {
    const int K = f.K, M = f.M, A = f.A;
    DeflArgs g;
    g.XX = XX; g.XY0 = XY; g.W = f.W; g.P = f.P; g.Q = f.Q; g.prof = prof; g.K = K; g.M = M; g.A = A;
    g.X = g.Y = nullptr; g.cube = nullptr; g.N = g.ldx = g.ldy = 0;
    g.c0 = 0; g.c1 = A; g.state = nullptr;
    int ldk = K;
    while (ldk % 16 != 4 && ldk % 16 != 12) ldk++;
    g.ldk = ldk;
    const size_t smem = defl_smem_doubles(K, M) * 8;
    double* U = ws_new<double>(ctx, (size_t)A * A);
    if (!U) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in pls_defl");
    const int KS = (K + 31) / 32;
    kernel_begin(ctx, 0);
#define DEFL_CASE(KS_)                                                                                                             \
    case KS_: {                                                                                                                    \
        auto kfn = pls_defl_kernel<KS_, false>;                                                                                    \
        CUDA_TRY(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                          \
        LAUNCH(ctx, kfn, 1, DT, smem, g);                                                                                          \
        break;                                                                                                                     \
    }
    switch (KS) {
        DEFL_CASE(1) DEFL_CASE(2) DEFL_CASE(3) DEFL_CASE(4) DEFL_CASE(5) DEFL_CASE(6)
        default: ABC_FAIL(ctx, ABCB200_EINVAL, "pls_defl: K=%d too large", K);
    }
#undef DEFL_CASE
    kernel_end(ctx, 0);
    return pls_ur_dev(ctx, f, U);
}

// true when the batched leave-one-out refits can run on chip for this shape
bool pls_loo_fits(const abcb200_ctx* ctx, int K, int M) {
    return K <= 192 && M <= 128 && defl_smem_doubles(K, M, true) * 8 + 512 <= (size_t)ctx->smem_optin;
}

// Model::cv_LOO (pls.cpp:469-491): cube[(y * A + c) * N + i] = residual of response y of row i predicted with c + 1 components by
// the model refitted without row i. XX = X^T X and XY = X^T Y over ALL N rows; one persistent CTA per SM walks the rows.
int pls_loo_dev(abcb200_ctx* ctx, const double* X, int64_t ldx, const double* Y, int64_t ldy, int64_t N, int K, int M, int A,
                const double* XX, const double* XY, double* cube) {
    DeflArgs g;
    g.XX = XX; g.XY0 = XY; g.W = g.P = g.Q = nullptr; g.prof = nullptr; g.K = K; g.M = M; g.A = A;
    g.X = X; g.Y = Y; g.cube = cube; g.N = N; g.ldx = ldx; g.ldy = ldy;
    g.c0 = 0; g.c1 = A; g.state = nullptr;
    int ldk = K;
    while (ldk % 16 != 4 && ldk % 16 != 12) ldk++;
    g.ldk = ldk;
    const size_t smem = defl_smem_doubles(K, M, true) * 8;
    const int grid = (int)std::min<int64_t>(N, ctx->sm_count);
    const int KS = (K + 31) / 32;
    kernel_begin(ctx, ABC_K_LOO);
#define LOO_CASE(KS_)                                                                                                              \
    case KS_: {                                                                                                                    \
        auto kfn = pls_defl_kernel<KS_, true>;                                                                                     \
        CUDA_TRY(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                          \
        LAUNCH(ctx, kfn, grid, DT, smem, g);                                                                                       \
        break;                                                                                                                     \
    }
    switch (KS) {
        LOO_CASE(1) LOO_CASE(2) LOO_CASE(3) LOO_CASE(4) LOO_CASE(5) LOO_CASE(6)
        default: ABC_FAIL(ctx, ABCB200_EINVAL, "pls_loo: K=%d too large", K);
    }
#undef LOO_CASE
    kernel_end(ctx, ABC_K_LOO);
    return ABCB200_OK;
}
