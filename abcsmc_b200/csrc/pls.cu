// pls.cu — S2 kernel-PLS fit and the dense FP64 products around it.
//
// Reference: PLS::Model::plsr, lib/PLS/src/pls.cpp:390-437 (Dayal & MacGregor 1997 modified kernel
// algorithms 1 and 2), Model::scores :439-442, Model::coefficients :444-447, ABC::euclidean
// src/AbcUtil.cpp:320-324.
//
// Device design
//  * XY = X^T Y (and X^T X for KERNEL_TYPE2) are tall-skinny FP64 contractions: DMMA (mma.sync m8n8k4 f64,
//    SASS DMMA.8x8x4) with fragments loaded straight from the column-major operands, split over row chunks,
//    partials reduced in a fixed order (deterministic). FP64 has no tcgen05 kind on sm_100a; DMMA and DFMA
//    share one pipe (profiles/r01_fp64_peak.json: 37.2 vs 33.9 TFLOP/s, 34.9 interleaved).
//  * Per component the reference does an M x M eigen-solve and O(K*M + K*i) vector work: one CTA
//    (pls_small_kernel). The dominant eigenvector of the symmetric PSD matrix XY^T XY is obtained by repeated
//    squaring (S <- S^2 / trace) to a rank-1 projector plus two power refinements with the original matrix:
//    log-depth, no sequential rotation chains, deterministic.
//  * KERNEL_TYPE1 (reference default) streams X once per component: pls_pass_kernel stages a row tile of X in
//    shared memory, computes t = X r for the tile, then p += X^T t and tt += t^T t from shared memory, so HBM
//    sees 8*n*(K+1) bytes per component. KERNEL_TYPE2 never re-reads X.
#include <stdlib.h>

#include "kernels.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// C += A^T B over a row chunk, DMMA. Warp tile (8*TA) x (8*TB) outputs.
// A-fragment (m = output row = column of A, k = data row): a = A[(ca + (lane>>2)) * lda + i + (lane&3)]
// B-fragment (k = data row, n = column of B):              b = B[(cb + (lane>>2)) * ldb + i + (lane&3)]
template <int TA, int TB>
__global__ void __launch_bounds__(256) atb_partial_kernel(const double* __restrict__ A, int64_t lda, int Ka,
                                                          const double* __restrict__ B, int64_t ldb, int Kb, int64_t n,
                                                          int64_t rows_per_chunk, double* __restrict__ partial) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int n_ta = (Ka + 8 * TA - 1) / (8 * TA), n_tb = (Kb + 8 * TB - 1) / (8 * TB);
    const int task = blockIdx.y * 8 + wid;
    if (task >= n_ta * n_tb) return;
    const int ca0 = (task % n_ta) * 8 * TA, cb0 = (task / n_ta) * 8 * TB;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_chunk;
    const int64_t r1 = min(n, r0 + rows_per_chunk);
    const int g = lane >> 2, q = lane & 3;

    const double* pa[TA]; bool va[TA];
    const double* pb[TB]; bool vb[TB];
#pragma unroll
    for (int x = 0; x < TA; x++) { const int c = ca0 + 8 * x + g; va[x] = c < Ka; pa[x] = A + (int64_t)min(c, Ka - 1) * lda + q; }
#pragma unroll
    for (int y = 0; y < TB; y++) { const int c = cb0 + 8 * y + g; vb[y] = c < Kb; pb[y] = B + (int64_t)min(c, Kb - 1) * ldb + q; }

    double acc[TA][TB][2];
#pragma unroll
    for (int x = 0; x < TA; x++)
#pragma unroll
        for (int y = 0; y < TB; y++) acc[x][y][0] = acc[x][y][1] = 0.0;

    int64_t i = r0;
    const int64_t r_full = r0 + ((r1 - r0) / 4) * 4;
#pragma unroll 4
    for (; i < r_full; i += 4) {
        double a[TA], b[TB];
#pragma unroll
        for (int x = 0; x < TA; x++) { const double v = pa[x][i]; a[x] = va[x] ? v : 0.0; }
#pragma unroll
        for (int y = 0; y < TB; y++) { const double v = pb[y][i]; b[y] = vb[y] ? v : 0.0; }
#pragma unroll
        for (int x = 0; x < TA; x++)
#pragma unroll
            for (int y = 0; y < TB; y++) dmma884(acc[x][y][0], acc[x][y][1], a[x], b[y]);
    }
    if (i < r1) {   // ragged tail: rows beyond r1 contribute zero
        const bool rv = (i + q) < r1;
        double a[TA], b[TB];
#pragma unroll
        for (int x = 0; x < TA; x++) a[x] = (va[x] && rv) ? pa[x][i] : 0.0;
#pragma unroll
        for (int y = 0; y < TB; y++) b[y] = (vb[y] && rv) ? pb[y][i] : 0.0;
#pragma unroll
        for (int x = 0; x < TA; x++)
#pragma unroll
            for (int y = 0; y < TB; y++) dmma884(acc[x][y][0], acc[x][y][1], a[x], b[y]);
    }
    double* out = partial + (int64_t)blockIdx.x * Ka * Kb;
#pragma unroll
    for (int x = 0; x < TA; x++)
#pragma unroll
        for (int y = 0; y < TB; y++) {
            const int row = ca0 + 8 * x + g;
            const int col = cb0 + 8 * y + 2 * q;
            if (row < Ka) {
                if (col < Kb) out[(int64_t)col * Ka + row] = acc[x][y][0];
                if (col + 1 < Kb) out[(int64_t)(col + 1) * Ka + row] = acc[x][y][1];
            }
        }
}

// out[j] = sum_c partial[c][j], c ascending (fixed order)
__global__ void reduce_partials_kernel(const double* __restrict__ partial, int nchunk, int64_t len, double* __restrict__ out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= len) return;
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};           // eight interleaved partial sums: loads in flight, order still fixed
    int c = 0;
    for (; c + 8 <= nchunk; c += 8) {
#pragma unroll
        for (int u = 0; u < 8; u++) a[u] += partial[(int64_t)(c + u) * len + j];
    }
#pragma unroll
    for (int u = 0; u < 8; u++) if (c + u < nchunk) a[u] += partial[(int64_t)(c + u) * len + j];
    out[j] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
}

// ------------------------------------------------------------------------------------------------
// The per-component scalar/vector work, one CTA.
struct SmallArgs {
    double* XY;            // K x M, ld K (deflated in place)
    const double* XX;      // K x K (KERNEL_TYPE2) or null
    double *W, *P, *R, *Q; // factor matrices
    const double* partial; // [npart][K+1] from pls_pass_kernel (KERNEL_TYPE1)
    int npart;
    double* r_cur;         // K: weights vector handed to pls_pass_kernel
    int K, M, A;
    int comp_begin, comp_end;   // components whose "start" phase runs in this launch
    int finish_prev;            // 1: first finish component comp_begin-1 from `partial`
};

constexpr int SMALL_THREADS = 512;

// trace of a padded symmetric matrix, computed redundantly by every warp (no block barrier needed)
__device__ __forceinline__ double warp_trace(const double* S, int M, int lds) {
    double t = 0;
    for (int a = threadIdx.x & 31; a < M; a += 32) t += S[a * lds + a];
    return warp_sum(t);
}

// Dominant eigenvector of the symmetric PSD M x M matrix S0 (shared memory, padded to Mp x Mp with row stride lds =
// Mp + 4, zero padding) by repeated squaring S <- S^2 / trace(S^2): the iterate converges quadratically to the
// projector q q^T. Each squaring is a DMMA product of 8x8 tiles spread over the warps (B fragments are read through
// the symmetry S[k][n] = S[n][k], which keeps shared-memory accesses conflict-free and the result bitwise symmetric).
// Result in qv (unit norm, largest-magnitude component positive) after two power refinements with S0.
__device__ void dominant_eigvec_squaring(const double* S0, double* S, double* S2, double* qv, int M, int Mp, int lds, double* red,
                                         int* ired) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int ntile = Mp / 8, ksteps = Mp / 4;
    const double tr = warp_trace(S0, M, lds);
    if (!(tr > 0.0)) {   // zero or NaN matrix: Eigen would hand back some unit vector; NaNs propagate downstream either way
        for (int a = tid; a < M; a += nt) qv[a] = (a == 0) ? 1.0 : 0.0;
        __syncthreads();
        return;
    }
    const double inv = 1.0 / tr;
    for (int i = tid; i < Mp * Mp; i += nt) { const int r = i / Mp, c = i - r * Mp; S[r * lds + c] = S0[r * lds + c] * inv; }
    __syncthreads();
    int extra = -1;
    for (int it = 0; it < 64; it++) {
        for (int tile = wid; tile < ntile * ntile; tile += nw) {
            const int ta = tile / ntile, tb = tile - ta * ntile;
            const double* pa = S + (ta * 8 + g) * lds + q;
            const double* pb = S + (tb * 8 + g) * lds + q;
            double c0 = 0.0, c1 = 0.0;
            for (int ks = 0; ks < ksteps; ks++) dmma884(c0, c1, pa[ks * 4], pb[ks * 4]);
            double* po = S2 + (ta * 8 + g) * lds + tb * 8 + 2 * q;
            po[0] = c0; po[1] = c1;
        }
        __syncthreads();
        const double inv2 = 1.0 / warp_trace(S2, M, lds);
        int conv = 1;
        for (int i = tid; i < Mp * Mp; i += nt) {
            const int r = i / Mp, c = i - r * Mp;
            const double v = S2[r * lds + c] * inv2;
            if (!(fabs(v - S[r * lds + c]) < 1e-8)) conv = 0;
            S[r * lds + c] = v;
        }
        const int all_conv = __syncthreads_and(conv);
        if (extra < 0) { if (all_conv) extra = 2; }      // a projector up to 1e-8: two more squarings -> 1e-32
        else if (--extra == 0) break;
    }
    if (tid == 0) {   // q = column of the projector with the largest diagonal entry
        int best = 0; double bv = S[0];
        for (int a = 1; a < M; a++) if (S[a * lds + a] > bv) { bv = S[a * lds + a]; best = a; }
        *ired = best;
    }
    __syncthreads();
    const int best = *ired;
    for (int a = tid; a < M; a += nt) qv[a] = S[a * lds + best];
    __syncthreads();
    for (int rep = 0; rep < 3; rep++) {   // rep 0: normalise; rep 1,2: power refinement with the original matrix
        if (rep > 0) {
            double v = 0;
            if (tid < M) { for (int l = 0; l < M; l++) v = fma(S0[tid * lds + l], qv[l], v); }
            __syncthreads();
            if (tid < M) qv[tid] = v;
            __syncthreads();
        }
        double nn = 0;
        for (int a = tid; a < M; a += nt) nn += qv[a] * qv[a];
        nn = block_sum(nn, red);
        const double sc = 1.0 / sqrt(nn);
        __syncthreads();
        for (int a = tid; a < M; a += nt) qv[a] *= sc;
        __syncthreads();
    }
    if (tid == 0) {
        int big = 0;
        for (int a = 1; a < M; a++) if (fabs(qv[a]) > fabs(qv[big])) big = a;
        *ired = (qv[big] < 0) ? 1 : 0;
    }
    __syncthreads();
    if (*ired) { for (int a = tid; a < M; a += nt) qv[a] = -qv[a]; }
    __syncthreads();
}

struct SmallSmem { double *S0, *S, *S2, *qv, *wv, *rv, *pv, *cv, *red; int* ired; double* s_tt; int Mp, lds; };

// start of component `comp` (pls.cpp:401-416): w (normalised), r; publishes W[:,comp], R[:,comp], r_cur
__device__ void pls_component_start(const SmallArgs& g, const SmallSmem& s, int comp) {
    const int K = g.K, M = g.M;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    if (M == 1) {                                                     // pls.cpp:403-404
        for (int k = tid; k < K; k += nt) s.wv[k] = g.XY[k];
        __syncthreads();
    } else {                                                          // pls.cpp:406-408
        // S0 = XY^T XY (pls.cpp:406) by DMMA: 8x8 tiles of the upper triangle, one per warp, mirrored (bitwise symmetric);
        // fragments straight from XY (column-major K x M): a = XY[k0 + (lane&3), ta*8 + (lane>>2)], b likewise with tb
        {
            const int Mp = s.Mp, lds = s.lds, ntile = Mp / 8;
            const int gq = lane >> 2, qq = lane & 3;
            const int npair = ntile * (ntile + 1) / 2;
            for (int pidx = wid; pidx < npair; pidx += nw) {
                int ta = 0, rem = pidx;
                while (rem >= ntile - ta) { rem -= ntile - ta; ta++; }
                const int tb = ta + rem;
                const int ca = ta * 8 + gq, cb = tb * 8 + gq;
                const bool va = ca < M, vb = cb < M;
                const double* pa = g.XY + (int64_t)min(ca, M - 1) * K;
                const double* pb = g.XY + (int64_t)min(cb, M - 1) * K;
                double c0 = 0.0, c1 = 0.0;
#pragma unroll 4
                for (int k0 = 0; k0 < K; k0 += 4) {
                    const int k = k0 + qq;
                    const bool kv = k < K;
                    const int kc = kv ? k : K - 1;
                    const double av = pa[kc], bv = pb[kc];
                    dmma884(c0, c1, (va && kv) ? av : 0.0, (vb && kv) ? bv : 0.0);
                }
                const int r = ta * 8 + gq, c = tb * 8 + 2 * qq;
                s.S0[r * lds + c] = c0; s.S0[r * lds + c + 1] = c1;
                if (ta != tb) { s.S0[c * lds + r] = c0; s.S0[(c + 1) * lds + r] = c1; }
            }
        }
        __syncthreads();
        dominant_eigvec_squaring(s.S0, s.S, s.S2, s.qv, M, s.Mp, s.lds, s.red, s.ired);
        for (int k = tid; k < K; k += nt) {
            double acc = 0;
            for (int m = 0; m < M; m++) acc = fma(g.XY[(int64_t)m * K + k], s.qv[m], acc);
            s.wv[k] = acc;
        }
        __syncthreads();
    }
    double ww = 0;
    for (int k = tid; k < K; k += nt) ww += s.wv[k] * s.wv[k];
    ww = block_sum(ww, s.red);
    const double wn = sqrt(ww);
    __syncthreads();
    for (int k = tid; k < K; k += nt) s.wv[k] /= wn;                  // pls.cpp:411
    __syncthreads();
    for (int j = wid; j < comp; j += 2 * nw) {                        // c_j = P_j^T w  (pls.cpp:415), two columns in flight
        const int j2 = j + nw;
        const double* pj = g.P + (int64_t)j * K;
        const double* pj2 = g.P + (int64_t)min(j2, comp - 1) * K;
        double acc = 0, acc2 = 0;
        for (int k = lane; k < K; k += 32) { const double wk = s.wv[k]; acc = fma(pj[k], wk, acc); acc2 = fma(pj2[k], wk, acc2); }
        acc = warp_sum(acc); acc2 = warp_sum(acc2);
        if (lane == 0) { s.cv[j] = acc; if (j2 < comp) s.cv[j2] = acc2; }
    }
    __syncthreads();
    for (int k = tid; k < K; k += nt) {                               // r = w - sum_j c_j R_j, j ascending
        double r = s.wv[k];
#pragma unroll 8
        for (int j = 0; j < comp; j++) r -= s.cv[j] * g.R[(int64_t)j * K + k];
        s.rv[k] = r;
        g.W[(int64_t)comp * K + k] = s.wv[k];
        g.R[(int64_t)comp * K + k] = r;
        g.r_cur[k] = r;
    }
    __syncthreads();
}

// finish component `comp` (pls.cpp:418-433): p, tt, q, deflate XY. Expects r in s.rv.
__device__ void pls_component_finish(const SmallArgs& g, const SmallSmem& s, int comp) {
    const int K = g.K, M = g.M;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    if (g.XX) {   // KERNEL_TYPE2 (pls.cpp:422-424): p = XX r, tt = r^T XX r
        for (int b = wid; b < K; b += nw) {
            const double* col = g.XX + (int64_t)b * K;
            double acc = 0;
            for (int a = lane; a < K; a += 32) acc = fma(s.rv[a], col[a], acc);
            acc = warp_sum(acc);
            if (lane == 0) s.pv[b] = acc;
        }
        __syncthreads();
        double t = 0;
        for (int k = tid; k < K; k += nt) t += s.pv[k] * s.rv[k];
        t = block_sum(t, s.red);
        if (tid == 0) *s.s_tt = t;
    } else {      // KERNEL_TYPE1 (pls.cpp:418-421): reduce the pass kernel's per-CTA partials in fixed order
        for (int k = tid; k <= K; k += nt) {
            double acc = 0;
#pragma unroll 8
            for (int c = 0; c < g.npart; c++) acc += g.partial[(int64_t)c * (K + 1) + k];
            if (k < K) s.pv[k] = acc; else *s.s_tt = acc;
        }
    }
    __syncthreads();
    const double tt = *s.s_tt;
    for (int k = tid; k < K; k += nt) { s.pv[k] /= tt; g.P[(int64_t)comp * K + k] = s.pv[k]; }   // pls.cpp:427
    __syncthreads();
    for (int m = wid; m < M; m += nw) {                                                          // pls.cpp:428
        const double* col = g.XY + (int64_t)m * K;
        double acc = 0;
        for (int k = lane; k < K; k += 32) acc = fma(s.rv[k], col[k], acc);
        acc = warp_sum(acc);
        if (lane == 0) { const double qm = acc / tt; s.qv[m] = qm; g.Q[(int64_t)comp * M + m] = qm; }
    }
    __syncthreads();
    for (int i = tid; i < K * M; i += nt) {                                                      // pls.cpp:429
        const int m = i / K, k = i - m * K;
        g.XY[i] -= (s.pv[k] * s.qv[m]) * tt;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(SMALL_THREADS) pls_small_kernel(SmallArgs g) {
    extern __shared__ double sm[];
    __shared__ int ired;
    __shared__ double s_tt;
    const int K = g.K, M = g.M;
    SmallSmem s;
    s.Mp = (M + 7) / 8 * 8; s.lds = s.Mp + 4;
    const int ssz = s.Mp * s.lds;
    s.S0 = sm; s.S = s.S0 + ssz; s.S2 = s.S + ssz; s.qv = s.S2 + ssz;
    s.wv = s.qv + s.Mp; s.rv = s.wv + K; s.pv = s.rv + K; s.cv = s.pv + K; s.red = s.cv + g.A;
    s.ired = &ired; s.s_tt = &s_tt;
    if (g.finish_prev) {
        const int prev = g.comp_begin - 1;
        for (int k = threadIdx.x; k < K; k += blockDim.x) s.rv[k] = g.R[(int64_t)prev * K + k];
        __syncthreads();
        pls_component_finish(g, s, prev);
    }
    for (int comp = g.comp_begin; comp < g.comp_end; comp++) {
        pls_component_start(g, s, comp);
        if (g.XX) pls_component_finish(g, s, comp);
    }
}

// ------------------------------------------------------------------------------------------------
// KERNEL_TYPE1 streaming pass for one component: t = X r, tt = t^T t, p = X^T t (un-normalised).
// Each CTA walks row tiles of RT rows; a tile (RT x K) is staged in shared memory once and used twice.
template <int RT>
__global__ void __launch_bounds__(256) pls_pass_kernel(const double* __restrict__ X, int64_t ld, int64_t n, int K,
                                                       const double* __restrict__ r, double* __restrict__ Tcol,
                                                       double* __restrict__ partial) {
    extern __shared__ double sm[];
    constexpr int G = 256 / RT;         // thread groups along k
    double* xs = sm;                    // K * RT, column k at xs + k*RT
    double* r_s = xs + (size_t)K * RT;  // K
    double* p_acc = r_s + K;            // K
    double* t_part = p_acc + K;         // G * RT
    double* t_s = t_part + G * RT;      // RT
    double* red = t_s + RT;             // 32
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int row = tid % RT, grp = tid / RT;
    for (int k = tid; k < K; k += 256) { r_s[k] = r[k]; p_acc[k] = 0.0; }
    double tt_acc = 0.0;
    __syncthreads();
    const int64_t ntiles = (n + RT - 1) / RT;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t i0 = tile * RT;
        const int64_t gi = i0 + row;
        const bool rv = gi < n;
        const double* xp = X + (rv ? gi : (n - 1));
        double acc = 0;
#pragma unroll 8
        for (int k = grp; k < K; k += G) {
            double v = xp[(int64_t)k * ld];
            v = rv ? v : 0.0;
            xs[(size_t)k * RT + row] = v;
            acc = fma(v, r_s[k], acc);
        }
        t_part[grp * RT + row] = acc;
        __syncthreads();
        if (tid < RT) {
            double t = 0;
#pragma unroll
            for (int gg = 0; gg < G; gg++) t += t_part[gg * RT + tid];
            t_s[tid] = t;
            if (Tcol && (i0 + tid) < n) Tcol[i0 + tid] = t;
            tt_acc = fma(t, t, tt_acc);
        }
        __syncthreads();
        for (int k = wid; k < K; k += 8) {
            const double* xc = xs + (size_t)k * RT;
            double s = 0;
#pragma unroll
            for (int rr = lane; rr < RT; rr += 32) s = fma(xc[rr], t_s[rr], s);
            s = warp_sum(s);
            if (lane == 0) p_acc[k] += s;
        }
        __syncthreads();
    }
    const double tt = block_sum(tt_acc, red);
    double* out = partial + (int64_t)blockIdx.x * (K + 1);
    for (int k = tid; k < K; k += 256) out[k] = p_acc[k];
    if (tid == 0) out[K] = tt;
}

// ------------------------------------------------------------------------------------------------
// KERNEL_TYPE1 streaming pass, TMA version (the one normally used): row tiles of RT rows x K columns are brought into
// an NSTAGE-deep shared-memory ring by bulk async copies (one cp.async.bulk per column segment, all arriving on the
// stage's mbarrier), so HBM reads of tile i+1.. overlap the arithmetic on tile i. Per tile: t = X_tile r (phase 1),
// then p += X_tile^T t with LANE-PRIVATE accumulators (phase 2) that are reduced across lanes once at the very end,
// so the inner loops are LDS + DFMA only. Requires 16-byte aligned columns (ld even, X 16-B aligned).
// The ragged last tile (n % RT rows) is loaded with ordinary predicated loads.
template <int RT, int CPW>   // CPW: columns per warp held in registers (0: accumulate through shared memory)
__global__ void __launch_bounds__(256) pls_pass_tma_kernel(const double* __restrict__ X, int64_t ld, int64_t n, int K,
                                                           const double* __restrict__ r, double* __restrict__ Tcol,
                                                           double* __restrict__ partial, int nstage) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int G = 256 / RT;
    const size_t stage_doubles = (size_t)K * RT;
    double* ring = (double*)smem_raw;
    double* r_s = ring + (size_t)nstage * stage_doubles;   // K
    double* p_acc = r_s + K;                               // K (only used when CPW == 0)
    double* t_part = p_acc + K;                            // G * RT
    double* t_s = t_part + G * RT;                         // RT
    double* red = t_s + RT;                                // 32
    uint64_t* full = (uint64_t*)(red + 32);                // nstage
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int row = tid % RT, grp = tid / RT;
    const int64_t nfull = n / RT;                          // full tiles
    const int rem = (int)(n - nfull * RT);                 // rows of the ragged tile (0: none)
    const int64_t my_full = (nfull > blockIdx.x) ? (nfull - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const bool my_tail = rem > 0 && (int)(nfull % gridDim.x) == (int)blockIdx.x;
    const uint32_t stage_bytes = (uint32_t)(stage_doubles * 8);

    if (tid == 0) {
        for (int s = 0; s < nstage; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int k = tid; k < K; k += 256) { r_s[k] = r[k]; p_acc[k] = 0.0; }
    __syncthreads();
    auto issue = [&](int64_t it) {   // warp 0 brings tile `it` of this CTA into slot it % nstage
        const int slot = (int)(it % nstage);
        const int64_t i0 = ((int64_t)blockIdx.x + it * gridDim.x) * RT;
        if (lane == 0) mbar_expect_tx(&full[slot], stage_bytes);
        __syncwarp();
        double* dst = ring + (size_t)slot * stage_doubles;
        for (int k = lane; k < K; k += 32) tma_bulk_g2s(dst + (size_t)k * RT, X + (int64_t)k * ld + i0, RT * 8, &full[slot]);
    };
    if (wid == 0) for (int64_t it = 0; it < nstage && it < my_full; it++) issue(it);

    double pacc[CPW > 0 ? CPW : 1];
#pragma unroll
    for (int j = 0; j < (CPW > 0 ? CPW : 1); j++) pacc[j] = 0.0;
    double tt_acc = 0.0;

    const int64_t ntl = my_full + (my_tail ? 1 : 0);
    for (int64_t it = 0; it < ntl; it++) {
        const bool tail = it >= my_full;
        const int slot = tail ? 0 : (int)(it % nstage);
        double* xs = ring + (size_t)slot * stage_doubles;
        int64_t i0;
        if (!tail) {
            i0 = ((int64_t)blockIdx.x + it * gridDim.x) * RT;
            mbar_wait(&full[slot], (uint32_t)((it / nstage) & 1));
        } else {   // ragged tile: every slot is drained by now; fill slot 0 by hand, zero rows beyond n
            i0 = nfull * RT;
            __syncthreads();
            const bool rv = row < rem;
            for (int k = grp; k < K; k += G) xs[(size_t)k * RT + row] = rv ? X[(int64_t)k * ld + i0 + row] : 0.0;
            __syncthreads();
        }
        // phase 1: t = X_tile r
        double acc = 0;
#pragma unroll 4
        for (int k = grp; k < K; k += G) acc = fma(xs[(size_t)k * RT + row], r_s[k], acc);
        t_part[grp * RT + row] = acc;
        __syncthreads();
        if (tid < RT) {
            double t = 0;
#pragma unroll
            for (int gg = 0; gg < G; gg++) t += t_part[gg * RT + tid];
            t_s[tid] = t;
            if (Tcol && (i0 + tid) < n) Tcol[i0 + tid] = t;
            tt_acc = fma(t, t, tt_acc);
        }
        __syncthreads();
        // phase 2: p += X_tile^T t
        if (CPW > 0) {
            double tl[(RT + 31) / 32];
#pragma unroll
            for (int j = 0; j < (RT + 31) / 32; j++) tl[j] = (lane + 32 * j < RT) ? t_s[lane + 32 * j] : 0.0;
#pragma unroll
            for (int j = 0; j < (CPW > 0 ? CPW : 1); j++) {
                const int k = wid + 8 * j;
                if (k < K) {
                    const double* xc = xs + (size_t)k * RT;
#pragma unroll
                    for (int jj = 0; jj < (RT + 31) / 32; jj++)
                        if (lane + 32 * jj < RT) pacc[j] = fma(xc[lane + 32 * jj], tl[jj], pacc[j]);
                }
            }
        } else {
            for (int k = wid; k < K; k += 8) {
                const double* xc = xs + (size_t)k * RT;
                double sacc = 0;
#pragma unroll
                for (int rr = lane; rr < RT; rr += 32) sacc = fma(xc[rr], t_s[rr], sacc);
                sacc = warp_sum(sacc);
                if (lane == 0) p_acc[k] += sacc;
            }
        }
        __syncthreads();   // every warp is done with this slot (and with t_s / t_part)
        if (wid == 0 && !tail && it + nstage < my_full) issue(it + nstage);
    }
    const double tt = block_sum(tt_acc, red);
    double* out = partial + (int64_t)blockIdx.x * (K + 1);
    if (CPW > 0) {
#pragma unroll
        for (int j = 0; j < (CPW > 0 ? CPW : 1); j++) {
            const int k = wid + 8 * j;
            const double v = warp_sum(pacc[j]);
            if (k < K && lane == 0) out[k] = v;
        }
    } else {
        for (int k = tid; k < K; k += 256) out[k] = p_acc[k];
    }
    if (tid == 0) out[K] = tt;
}

// ------------------------------------------------------------------------------------------------
// C = X * B (DMMA). MODE 0: store C (n x ncols). MODE 1: dist[i] = || C[i,:] - ref ||_2 (no C in HBM).
// Warp tile: 32 rows (4 m-tiles) x 32 columns (4 n-tiles); a warp walks all column groups of its rows.
// A-fragment (m = data row, k): a = X[(k0 + (lane&3)) * ldx + row0 + 8x + (lane>>2)]
// B-fragment (k, n = column):   b = B[(c0 + 8y + (lane>>2)) * ldb + k0 + (lane&3)]
#ifndef XB_PF_STEPS
#define XB_PF_STEPS 4
#endif
constexpr int XB_PF = XB_PF_STEPS;   // L2 prefetch distance of xb_kernel's A operand, in k-steps of 4 columns
constexpr int XB_WARPS = 4;      // 128-thread CTAs, three per SM: the DMMA pipe wants >= 3 ready warps per scheduler (a warp issues one DMMA per ~26 cycles)
template <int MODE>
__global__ void __launch_bounds__(32 * XB_WARPS, 3) xb_kernel(const double* __restrict__ X, int64_t ldx, int64_t n, int K,
                                                 const double* __restrict__ B, int64_t ldb, int ncols,
                                                 const double* __restrict__ ref, double* __restrict__ out, int64_t ldo,
                                                 int64_t split, int64_t gap, int pf_steps) {
    // MODE 0: output row of data row r is r + (r >= split ? gap : 0): the pipelined ranking keeps the hold-out rows of the score
    // matrix on a 256-byte boundary of their own (the selection kernels stream them) although the training rows above them end anywhere.
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int64_t nblk = (n + 31) / 32;
    // MODE 0: a work unit is (32 rows, 32 output columns), the column groups of a row block being neighbours (they share
    // the rows through L1), so short hold-out sets still fill the machine evenly; MODE 1 keeps a row block's groups in one
    // warp (the squared distance accumulates over them).
    const int ngrp = (MODE == 0) ? (ncols + 31) / 32 : 1;
    const int64_t nunit = nblk * ngrp;
    for (int64_t unit = (int64_t)blockIdx.x * XB_WARPS + wid; unit < nunit; unit += (int64_t)gridDim.x * XB_WARPS) {
        const int64_t blk = unit / ngrp;
        const int cbeg = (MODE == 0) ? (int)(unit % ngrp) * 32 : 0;
        const int cend = (MODE == 0) ? min(cbeg + 32, ncols) : ncols;
        const int64_t row0 = blk * 32;
        const double* pa[4]; bool va[4];
#pragma unroll
        for (int x = 0; x < 4; x++) { const int64_t rr = row0 + 8 * x + g; va[x] = rr < n; pa[x] = X + min(rr, n - 1); }
        double rowacc[4] = {0, 0, 0, 0};
        for (int c0 = cbeg; c0 < cend; c0 += 32) {
            const double* pb[4]; bool vb[4];
#pragma unroll
            for (int y = 0; y < 4; y++) { const int c = c0 + 8 * y + g; vb[y] = c < ncols; pb[y] = B + (int64_t)min(c, ncols - 1) * ldb; }
            double acc[4][4][2];
#pragma unroll
            for (int x = 0; x < 4; x++)
#pragma unroll
                for (int y = 0; y < 4; y++) acc[x][y][0] = acc[x][y][1] = 0.0;
#pragma unroll 2
            for (int k0 = 0; k0 < K; k0 += 4) {
                const int k = k0 + q;
                const bool kv = k < K;
                const int kc = kv ? k : (K - 1);
                // the A fragments come straight from the column-major operand in HBM and the warps wait on them (62 % of the stall
                // samples, ncu r02): the two 128-byte lines of each of the four columns XB_PF steps ahead are pulled into L2
                if (g == 0 && k + 4 * pf_steps < K) {
                    const double* pf = X + (int64_t)(k + 4 * pf_steps) * ldx + min(row0, n - 1);
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + min((int64_t)16, n - 1 - min(row0, n - 1))));
                }
                double a[4], b[4];
#pragma unroll
                for (int x = 0; x < 4; x++) { const double v = pa[x][(int64_t)kc * ldx]; a[x] = (va[x] && kv) ? v : 0.0; }
#pragma unroll
                for (int y = 0; y < 4; y++) { const double v = pb[y][kc]; b[y] = (vb[y] && kv) ? v : 0.0; }
#pragma unroll
                for (int x = 0; x < 4; x++)
#pragma unroll
                    for (int y = 0; y < 4; y++) dmma884(acc[x][y][0], acc[x][y][1], a[x], b[y]);
            }
#pragma unroll
            for (int x = 0; x < 4; x++)
#pragma unroll
                for (int y = 0; y < 4; y++) {
                    const int col = c0 + 8 * y + 2 * q;
                    const int64_t rr = row0 + 8 * x + g;
                    if (MODE == 0) {
                        if (rr < n) {
                            const int64_t ro = rr + (rr >= split ? gap : 0);
                            if (col < ncols) out[(int64_t)col * ldo + ro] = acc[x][y][0];
                            if (col + 1 < ncols) out[(int64_t)(col + 1) * ldo + ro] = acc[x][y][1];
                        }
                    } else {
                        if (col < ncols) { const double d = acc[x][y][0] - ref[col]; rowacc[x] = fma(d, d, rowacc[x]); }
                        if (col + 1 < ncols) { const double d = acc[x][y][1] - ref[col + 1]; rowacc[x] = fma(d, d, rowacc[x]); }
                    }
                }
        }
        if (MODE == 1) {
#pragma unroll
            for (int x = 0; x < 4; x++) {
                double v = rowacc[x];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                const int64_t rr = row0 + 8 * x + g;
                if (q == 0 && rr < n) out[rr] = sqrt(v);
            }
        }
    }
}

__global__ void vec_times_mat_kernel(const double* __restrict__ v, int K, const double* __restrict__ B, int64_t ldb, int ncols,
                                     double* __restrict__ out) {
    const int a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (a >= ncols) return;
    const int lane = threadIdx.x & 31;
    double s = 0;
    for (int k = lane; k < K; k += 32) s = fma(v[k], B[(int64_t)a * ldb + k], s);
    s = warp_sum(s);
    if (lane == 0) out[a] = s;
}

// C[k, m] = sum_{a < comp} R[k, a] Q[m, a]   (pls.cpp:444-447)
__global__ void coefficients_kernel(const double* __restrict__ R, const double* __restrict__ Q, int K, int M, int comp,
                                    double* __restrict__ C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * M) return;
    const int m = i / K, k = i - m * K;
    double s = 0;
    for (int a = 0; a < comp; a++) s = fma(R[(int64_t)a * K + k], Q[(int64_t)a * M + m], s);
    C[i] = s;
}

size_t small_smem_bytes(int K, int M, int A) {
    const size_t Mp = (size_t)(M + 7) / 8 * 8;
    return sizeof(double) * (3 * Mp * (Mp + 4) + Mp + 3 * (size_t)K + A + 32);
}

int pass_rt(const abcb200_ctx* ctx, int K) {
    const size_t budget = (size_t)ctx->smem_optin - 1024;
    const int cand[4] = {256, 128, 64, 32};
    for (int c = 0; c < 4; c++) {
        const int RT = cand[c];
        const size_t need = sizeof(double) * ((size_t)K * RT + 2 * (size_t)K + (256 / RT) * RT + RT + 32);
        if (need <= budget && (RT <= 128 || K <= 64)) return RT;
    }
    return 0;
}
size_t pass_smem_bytes(int K, int RT) { return sizeof(double) * ((size_t)K * RT + 2 * (size_t)K + (256 / RT) * RT + RT + 32); }

void atb_plan(const abcb200_ctx* ctx, int64_t n, int Ka, int Kb, int* nchunk, int64_t* rows_per_chunk, int* gy) {
    const int n_ta = (Ka + 15) / 16, n_tb = (Kb + 15) / 16;
    *gy = (n_ta * n_tb + 7) / 8;
    int64_t want = (int64_t)(4 * ctx->sm_count) / (*gy);
    if (want < 1) want = 1;
    int64_t rpc = (n + want - 1) / want;
    if (rpc < 256) rpc = 256;
    rpc = (rpc + 3) / 4 * 4;
    *rows_per_chunk = rpc;
    *nchunk = (int)((n + rpc - 1) / rpc);
}

}  // namespace

size_t atb_ws_bytes(const abcb200_ctx* ctx, int64_t n, int Ka, int Kb) {
    int nchunk, gy; int64_t rpc;
    atb_plan(ctx, n, Ka, Kb, &nchunk, &rpc, &gy);
    return align_up((size_t)nchunk * Ka * Kb * sizeof(double), 256) + 256;
}

int launch_atb(abcb200_ctx* ctx, const double* A, int64_t lda, int Ka, const double* B, int64_t ldb, int Kb, int64_t n, double* C) {
    int nchunk, gy; int64_t rpc;
    atb_plan(ctx, n, Ka, Kb, &nchunk, &rpc, &gy);
    double* partial = ws_new<double>(ctx, (size_t)nchunk * Ka * Kb);
    if (!partial) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in launch_atb");
    LAUNCH(ctx, (atb_partial_kernel<2, 2>), dim3(nchunk, gy), 256, 0, A, lda, Ka, B, ldb, Kb, n, rpc, partial);
    const int64_t len = (int64_t)Ka * Kb;
    LAUNCH(ctx, reduce_partials_kernel, (unsigned)((len + 255) / 256), 256, 0, partial, nchunk, len, C);
    return ABCB200_OK;
}

size_t pls_fit_ws_bytes(const abcb200_ctx* ctx, int64_t n, int K, int M, int method) {
    size_t b = 0;
    b += align_up((size_t)K * M * 8, 256);                       // XY
    b += atb_ws_bytes(ctx, n, K, M);
    b += pls_gram_ws_bytes(ctx, n, K, M);
    b += align_up((size_t)K * 8, 256);                           // r_cur
    b += align_up((size_t)(2 * ctx->sm_count) * (K + 1) * 8, 256);   // pass partials
    return b + 4096;
}

int pls_fit_dev(abcb200_ctx* ctx, const double* X, int64_t ldx, const double* Y, int64_t ldy, PlsFactors f) {
    const int K = f.K, M = f.M, A = f.A;
    const int64_t n = f.n;
    if (K < 1 || M < 1 || A < 1 || A > K || n < 1) ABC_FAIL(ctx, ABCB200_EINVAL, "pls_fit: bad shape n=%lld K=%d M=%d A=%d", (long long)n, K, M, A);
    if (M > 128) ABC_FAIL(ctx, ABCB200_EINVAL, "pls_fit: M=%d responses exceed the on-chip eigen-solver limit (128)", M);
    const size_t ssm = small_smem_bytes(K, M, A);
    if (ssm > (size_t)ctx->smem_optin) ABC_FAIL(ctx, ABCB200_EINVAL, "pls_fit: K=%d M=%d A=%d need %zu B of shared memory (> %d)", K, M, A, ssm, ctx->smem_optin);
    CUDA_TRY(ctx, cudaFuncSetAttribute(pls_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm));

    if (f.method != ABCB200_KERNEL_TYPE1_STREAM) {
        // KERNEL_TYPE1 and KERNEL_TYPE2 share the Gram-based component loop (pls_gram.cu); they differ in whether the
        // training scores T = X R are materialised (pls.cpp:394, 418, 434).
        ABC_TRY(pls_fit_gram_dev(ctx, X, ldx, Y, ldy, f));
        if (f.T) ABC_TRY(launch_xb(ctx, X, ldx, n, K, f.R, K, A, f.T, f.ldt));
        return ABCB200_OK;
    }
    double* XY = ws_new<double>(ctx, (size_t)K * M);
    double* r_cur = ws_new<double>(ctx, K);
    if (!XY || !r_cur) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in pls_fit");
    ABC_TRY(launch_atb(ctx, X, ldx, K, Y, ldy, M, n, XY));        // pls.cpp:396

    SmallArgs g;
    g.XY = XY; g.XX = nullptr; g.W = f.W; g.P = f.P; g.R = f.R; g.Q = f.Q; g.partial = nullptr; g.npart = 0;
    g.r_cur = r_cur; g.K = K; g.M = M; g.A = A;

    // ---- streaming pass plan: TMA ring when the columns are 16-byte aligned, plain loads otherwise ------------------
    const bool aligned = (ldx % 2 == 0) && (((uintptr_t)X) % 16 == 0);
    int RT = 0, nstage = 0;
    const size_t budget = (size_t)ctx->smem_optin - 2048;
    if (aligned) {
        const int cand[4] = {128, 64, 32, 16};
        for (int c = 0; c < 4 && RT == 0; c++) {
            const size_t fixed = sizeof(double) * (2 * (size_t)K + 256 + cand[c] + 32) + 8 * 8 + 128;
            const size_t stage = (size_t)K * cand[c] * 8;
            if (budget < fixed) break;
            const int ns = (int)((budget - fixed) / stage);
            if (ns >= 3 || (cand[c] == 16 && ns >= 2)) { RT = cand[c]; nstage = ns > 8 ? 8 : ns; }
        }
    }
    if (RT > 0) {
        const size_t psm = (size_t)nstage * K * RT * 8 + sizeof(double) * (2 * (size_t)K + 256 + RT + 32) + 8 * 8 + 128;
        const bool regacc = (K + 7) / 8 <= 20;
        const int64_t ntiles = (n + RT - 1) / RT;
        const int grid = (int)min(ntiles, (int64_t)ctx->sm_count);
        double* partial = ws_new<double>(ctx, (size_t)grid * (K + 1));
        if (!partial) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in pls_fit");
        g.partial = partial; g.npart = grid;
#define PASS_DISPATCH(RTV)                                                                                                          \
    if (regacc) {                                                                                                                   \
        if (comp == 0) CUDA_TRY(ctx, cudaFuncSetAttribute(pls_pass_tma_kernel<RTV, 20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm)); \
        LAUNCH(ctx, (pls_pass_tma_kernel<RTV, 20>), grid, 256, psm, X, ldx, n, K, r_cur, Tcol, partial, nstage);                    \
    } else {                                                                                                                        \
        if (comp == 0) CUDA_TRY(ctx, cudaFuncSetAttribute(pls_pass_tma_kernel<RTV, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));  \
        LAUNCH(ctx, (pls_pass_tma_kernel<RTV, 0>), grid, 256, psm, X, ldx, n, K, r_cur, Tcol, partial, nstage);                     \
    }
        for (int comp = 0; comp <= A; comp++) {
            g.comp_begin = comp; g.comp_end = (comp < A) ? comp + 1 : comp; g.finish_prev = comp > 0 ? 1 : 0;
            LAUNCH(ctx, pls_small_kernel, 1, SMALL_THREADS, ssm, g);
            if (comp == A) break;
            double* Tcol = f.T ? f.T + (int64_t)comp * f.ldt : nullptr;
            switch (RT) {
                case 128: PASS_DISPATCH(128) break;
                case 64: PASS_DISPATCH(64) break;
                case 32: PASS_DISPATCH(32) break;
                default: PASS_DISPATCH(16) break;
            }
        }
#undef PASS_DISPATCH
        return ABCB200_OK;
    }
    // ---- fallback: unaligned operands ----------------------------------------------------------------------------------
    RT = pass_rt(ctx, K);
    if (RT == 0) ABC_FAIL(ctx, ABCB200_EINVAL, "pls_fit: K=%d too wide for the KERNEL_TYPE1 streaming tile; use KERNEL_TYPE2", K);
    const size_t psm = pass_smem_bytes(K, RT);
    const int64_t ntiles = (n + RT - 1) / RT;
    int occ = (int)((size_t)ctx->smem_optin / (psm + 1024)); if (occ < 1) occ = 1; if (occ > 2) occ = 2;
    int grid = (int)min(ntiles, (int64_t)occ * ctx->sm_count);
    double* partial = ws_new<double>(ctx, (size_t)grid * (K + 1));
    if (!partial) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in pls_fit");
    g.partial = partial; g.npart = grid;
    switch (RT) {
        case 256: CUDA_TRY(ctx, cudaFuncSetAttribute(pls_pass_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm)); break;
        case 128: CUDA_TRY(ctx, cudaFuncSetAttribute(pls_pass_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm)); break;
        case 64: CUDA_TRY(ctx, cudaFuncSetAttribute(pls_pass_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm)); break;
        default: CUDA_TRY(ctx, cudaFuncSetAttribute(pls_pass_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm)); break;
    }
    for (int comp = 0; comp <= A; comp++) {
        g.comp_begin = comp; g.comp_end = (comp < A) ? comp + 1 : comp; g.finish_prev = comp > 0 ? 1 : 0;
        LAUNCH(ctx, pls_small_kernel, 1, SMALL_THREADS, ssm, g);
        if (comp == A) break;
        double* Tcol = f.T ? f.T + (int64_t)comp * f.ldt : nullptr;
        switch (RT) {
            case 256: LAUNCH(ctx, pls_pass_kernel<256>, grid, 256, psm, X, ldx, n, K, r_cur, Tcol, partial); break;
            case 128: LAUNCH(ctx, pls_pass_kernel<128>, grid, 256, psm, X, ldx, n, K, r_cur, Tcol, partial); break;
            case 64: LAUNCH(ctx, pls_pass_kernel<64>, grid, 256, psm, X, ldx, n, K, r_cur, Tcol, partial); break;
            default: LAUNCH(ctx, pls_pass_kernel<32>, grid, 256, psm, X, ldx, n, K, r_cur, Tcol, partial); break;
        }
    }
    return ABCB200_OK;
}

static int xb_pf() { static const int v = getenv("ABCB200_XB_PF") ? atoi(getenv("ABCB200_XB_PF")) : XB_PF; return v; }      // tuning knob

int launch_xb(abcb200_ctx* ctx, const double* X, int64_t ldx, int64_t n, int K, const double* B, int64_t ldb, int ncols,
              double* out, int64_t ldo, int64_t split, int64_t gap) {
    if (n <= 0 || ncols <= 0) return ABCB200_OK;
    const int64_t nblk = (n + 31) / 32;
    const int64_t nunit = nblk * ((ncols + 31) / 32);
    int grid = (int)min((nunit + XB_WARPS - 1) / XB_WARPS, (int64_t)(12 * ctx->sm_count));
    LAUNCH(ctx, xb_kernel<0>, grid, 32 * XB_WARPS, 0, X, ldx, n, K, B, ldb, ncols, (const double*)nullptr, out, ldo, split, gap, xb_pf());
    return ABCB200_OK;
}

int launch_project_dist(abcb200_ctx* ctx, const double* X, int64_t ldx, int64_t n, int K, const double* B, int64_t ldb, int ncols,
                        const double* ref_scores, double* dist) {
    if (n <= 0) return ABCB200_OK;
    const int64_t nblk = (n + 31) / 32;
    int grid = (int)min((nblk + XB_WARPS - 1) / XB_WARPS, (int64_t)(12 * ctx->sm_count));
    LAUNCH(ctx, xb_kernel<1>, grid, 32 * XB_WARPS, 0, X, ldx, n, K, B, ldb, ncols, ref_scores, dist, (int64_t)0, (int64_t)0, (int64_t)0, xb_pf());
    return ABCB200_OK;
}

int launch_vec_times_mat(abcb200_ctx* ctx, const double* v, int K, const double* B, int64_t ldb, int ncols, double* out) {
    LAUNCH(ctx, vec_times_mat_kernel, (ncols + 3) / 4, 128, 0, v, K, B, ldb, ncols, out);
    return ABCB200_OK;
}

int launch_coefficients(abcb200_ctx* ctx, const double* R, const double* Q, int K, int M, int comp, double* C) {
    LAUNCH(ctx, coefficients_kernel, (K * M + 255) / 256, 256, 0, R, Q, K, M, comp, C);
    return ABCB200_OK;
}
