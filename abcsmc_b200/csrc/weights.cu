// weights.cu — S8 SMC importance-weight update, the O(N_new x N_old x P) hot loop.
//
// Reference: ABC::weight_predictive_prior, src/AbcUtil.cpp:547-586:
//   den_i = sum_j w_j prod_p phi(theta_new[i,p] - theta_old[j,p]; sigma_p = sqrt(dv_p)),  weight_i = numer_i / den_i,
//   then Eigen normalize() (L2).  phi = gsl_ran_gaussian_pdf. A factor p is skipped iff dv_p == 0 and the two
//   values are equal (:573); dv_p == 0 with different values gives inf * 0 = NaN for that row.
//
// Device design (FP64 pipe bound; DMMA and DFMA share that pipe on B200, profiles/r01_fp64_peak.json):
//   den_i = C * sum_j exp(-(|a_i|^2 + |b_j|^2 - 2 a_i.b_j - ln w_j)),  a = (theta_new - c) / sqrt(2 dv), b likewise,
//   C = prod_p 1/(sqrt(2 pi) sigma_p), c = first old particle (centring keeps |a|,|b| small so the expanded form
//   loses < 1e-13 in the exponent; a conditioning check falls back to the pairwise-difference kernel otherwise).
//   The whole exponent is ONE inner product of augmented vectors a' = [-2a, |a|^2, 1], b' = [b, 1, |b|^2 - ln w],
//   evaluated by DMMA.8x8x4 on fragments: new particles live in registers for the CTA's lifetime, old particles are
//   pre-packed into fragment order so a stage is one contiguous block fetched by TMA bulk copy (cp.async.bulk +
//   mbarrier, SASS UBLKCP) into a 4-stage shared-memory ring. The epilogue is a hand-rolled FP64 exp (no SFU path
//   for FP64). Algorithmic work per pair: 3P+2 flops + 1 exp.
#include "kernels.cuh"

namespace {

constexpr int W_THREADS = 256;
constexpr int W_TJ = 64;          // old particles per pipeline stage
constexpr int W_STAGES = 4;

// scale[p] = 1/sqrt(2 dv_p) (0 when dv_p == 0: the dimension leaves the quadratic form), centre[p] = theta_old[0,p],
// cinv = prod_{dv_p != 0} sqrt(2 pi dv_p) (1/C); for dv_p == 0 also the old column's min and max.
__global__ void weights_prep_kernel(const double* __restrict__ th_old, int64_t ld_old, int64_t n_old, const double* __restrict__ dv,
                                    int P, double* __restrict__ scale, double* __restrict__ centre, double* __restrict__ cinv,
                                    double* __restrict__ colmin, double* __restrict__ colmax, int* __restrict__ poison) {
    __shared__ double red[32];
    const int p = blockIdx.x;
    const double d = dv[p];
    if (threadIdx.x == 0) {
        scale[p] = (d == 0.0) ? 0.0 : 1.0 / sqrt(2.0 * d);
        centre[p] = th_old[(int64_t)p * ld_old];
        if (d != d) *poison = 1;
    }
    if (d == 0.0) {
        double mn = 1.0 / 0.0, mx = -1.0 / 0.0;
        for (int64_t j = threadIdx.x; j < n_old; j += blockDim.x) { const double v = th_old[(int64_t)p * ld_old + j]; mn = fmin(mn, v); mx = fmax(mx, v); }
        mn = -warp_max(-mn); mx = warp_max(mx);
        if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = mn; red[8 + (threadIdx.x >> 5)] = mx; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < (int)(blockDim.x >> 5); w++) { mn = fmin(mn, red[w]); mx = fmax(mx, red[8 + w]); }
            colmin[p] = mn; colmax[p] = mx;
        }
    }
    if (p == 0 && threadIdx.x == 0) {
        double c = 1.0;
        for (int k = 0; k < P; k++) { const double s = sqrt(dv[k]); if (dv[k] != 0.0) c *= sqrt(2.0 * M_PI) * fabs(s); }   // 1/C, gsl formula
        *cinv = c;
        cinv[3] = 4.0 + 0.25 * (double)(P + 2);      // rounding terms of the expanded exponent, see gate_says_dmma (cinv = scal)
    }
}

// packed[(jt*KS + s)*32 + (j&7)*4 + (k&3)], k = 4s + (k&3), KS k-steps; old particles:
//   k < P: (theta - c)*scale; k == P: 1; k == P+1: |b|^2 - ln w; rest 0. Padding particles: exponent slot 1e300.
__global__ void pack_old_kernel(const double* __restrict__ th, int64_t ld, int64_t n, int64_t n_pad, const double* __restrict__ w,
                                const double* __restrict__ scale, const double* __restrict__ centre, int P, int KS,
                                double* __restrict__ packed, double* __restrict__ maxnorm, int* __restrict__ poison) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_pad) return;
    double* dst = packed + (j >> 3) * (int64_t)KS * 32 + (j & 7) * 4;
    const bool valid = j < n;
    double nb = 0;
    bool bad = false;
    const int kend = max(4 * KS, P + 2);
    for (int k = 0; k < kend; k++) {
        double v = 0.0;
        if (valid && k < P) {
            const double t = th[(int64_t)k * ld + j];
            bad |= (t != t);
            v = (t - centre[k]) * scale[k];
            nb = fma(v, v, nb);
        } else if (k == P) {
            v = 1.0;
        } else if (k == P + 1) {
            if (valid) { const double wj = w[j]; bad |= (wj != wj); v = nb - log(wj); } else v = 1e300;
        }
        if (k < 4 * KS) dst[(k >> 2) * 32 + (k & 3)] = v;
    }
    if (bad) *poison = 1;
    if (valid) atomicMax((unsigned long long*)maxnorm, (unsigned long long)__double_as_longlong(nb));   // nb >= 0: bit order == value order
}

// new particles: k < P: -2 (theta - c)*scale; k == P: |a|^2; k == P+1: 1. Also the per-row NaN flag (:573 edge case).
__global__ void pack_new_kernel(const double* __restrict__ th, int64_t ld, int64_t n, int64_t n_pad, const double* __restrict__ scale,
                                const double* __restrict__ centre, const double* __restrict__ dv, const double* __restrict__ colmin,
                                const double* __restrict__ colmax, int P, int KS, double* __restrict__ packed,
                                double* __restrict__ maxnorm, int* __restrict__ nanflag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    double* dst = packed + (i >> 3) * (int64_t)KS * 32 + (i & 7) * 4;
    const bool valid = i < n;
    double na = 0;
    bool bad = false;
    const int kend = max(4 * KS, P + 2);
    for (int k = 0; k < kend; k++) {
        double v = 0.0;
        if (valid && k < P) {
            const double t = th[(int64_t)k * ld + i];
            bad |= (t != t);
            if (dv[k] == 0.0) bad |= (t != colmin[k]) || (t != colmax[k]);   // converged parameter but a differing value: inf*0
            const double a = (t - centre[k]) * scale[k];
            na = fma(a, a, na);
            v = -2.0 * a;
        } else if (k == P) {
            v = valid ? na : 0.0;
        } else if (k == P + 1) {
            v = valid ? 1.0 : 0.0;
        }
        if (k < 4 * KS) dst[(k >> 2) * 32 + (k & 3)] = v;
    }
    if (valid) {
        nanflag[i] = bad ? 1 : 0;
        atomicMax((unsigned long long*)maxnorm, (unsigned long long)__double_as_longlong(na));
    }
}

// Device-side choice between the two formulations in auto mode (no host round trip): the expanded exponent loses about
// (4 + (P + 2) / 4) eps (max |a|^2 + max |b|^2) absolutely (the operands' own roundings plus the P + 2 products accumulated by
// DMMA, |a_k b_k| <= (a_k^2 + b_k^2) / 2); the DMMA kernel runs when that is below 1e-12, the pairwise-difference kernel
// otherwise. Both are launched, the one whose turn it is not returns at once.
// gate = scal ([1] max |a|^2 over the new rows — of ALL shards when the update is sharded, sharded.cu —, [2] max |b|^2, [3] the factor) or null.
__device__ __forceinline__ bool gate_says_dmma(const double* __restrict__ gate) {
    const double bound = gate[3] * 2.220446049250313e-16 * (gate[1] + gate[2]);
    return bound < 1e-12;
}

// ------------------------------------------------------------------------------------------------
// Main kernel. grid = (row blocks, j splits). Each warp owns NI i-tiles (8 rows each) whose A fragments stay in
// registers; the CTA streams its share of packed old particles through a W_STAGES-deep TMA ring.
template <int KS, int NI>
__global__ void __launch_bounds__(W_THREADS) weights_dmma_kernel(const double* __restrict__ Apk, const double* __restrict__ Bpk,
                                                                 int64_t n_new_pad, int64_t n_old_pad, int stages_per_split,
                                                                 double* __restrict__ den_part, const double* __restrict__ gate) {
    if (gate && !gate_says_dmma(gate)) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int STAGE_DOUBLES = (W_TJ / 8) * KS * 32;
    constexpr uint32_t STAGE_BYTES = STAGE_DOUBLES * 8;
    double* ring = (double*)smem_raw;
    uint64_t* full = (uint64_t*)(smem_raw + (size_t)W_STAGES * STAGE_BYTES);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t total_stages = n_old_pad / W_TJ;
    const int64_t st0 = (int64_t)blockIdx.y * stages_per_split;
    const int64_t st1 = min(total_stages, st0 + stages_per_split);
    const int nst = (int)(st1 - st0);

    if (tid == 0) {
        for (int s = 0; s < W_STAGES; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < W_STAGES && s < nst; s++) {
            mbar_expect_tx(&full[s], STAGE_BYTES);
            tma_bulk_g2s(ring + (size_t)s * STAGE_DOUBLES, Bpk + (st0 + s) * (int64_t)STAGE_DOUBLES, STAGE_BYTES, &full[s]);
        }
    }
    // A fragments: rows of this warp, resident in registers
    const int64_t it0 = ((int64_t)blockIdx.x * (W_THREADS / 32) + wid) * NI;
    double af[NI][KS];
#pragma unroll
    for (int x = 0; x < NI; x++)
#pragma unroll
        for (int s = 0; s < KS; s++) af[x][s] = Apk[((it0 + x) * KS + s) * 32 + lane];
    double den[NI];
#pragma unroll
    for (int x = 0; x < NI; x++) den[x] = 0.0;

    for (int st = 0; st < nst; st++) {
        const int slot = st % W_STAGES;
        mbar_wait(&full[slot], (uint32_t)((st / W_STAGES) & 1));
        const double* bs = ring + (size_t)slot * STAGE_DOUBLES;
#pragma unroll 2
        for (int jb = 0; jb < W_TJ / 8; jb++) {
            double bf[KS];
#pragma unroll
            for (int s = 0; s < KS; s++) bf[s] = bs[(jb * KS + s) * 32 + lane];
#pragma unroll
            for (int x = 0; x < NI; x++) {
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int s = 0; s < KS; s++) dmma884(c0, c1, af[x][s], bf[s]);
                den[x] += exp_neg(c0) + exp_neg(c1);
            }
        }
        __syncthreads();   // every warp is done with this slot
        if (tid == 0 && st + W_STAGES < nst) {
            mbar_expect_tx(&full[slot], STAGE_BYTES);
            tma_bulk_g2s(ring + (size_t)slot * STAGE_DOUBLES, Bpk + (st0 + st + W_STAGES) * (int64_t)STAGE_DOUBLES, STAGE_BYTES, &full[slot]);
        }
    }
#pragma unroll
    for (int x = 0; x < NI; x++) {
        double v = den[x];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        const int64_t row = (it0 + x) * 8 + (lane >> 2);
        if ((lane & 3) == 0 && row < n_new_pad) den_part[(int64_t)blockIdx.y * n_new_pad + row] = v;
    }
}

// Pairwise-difference kernel (the reference's formulation; any P; used when the expanded form is ill-conditioned).
// CTA = 128 rows; a-tile [p][128] and a b-tile of 32 old particles [j][p] in shared memory.
__global__ void __launch_bounds__(128) weights_diff_kernel(const double* __restrict__ th_new, int64_t ld_new, int64_t n_new,
                                                           const double* __restrict__ th_old, int64_t ld_old, int64_t n_old,
                                                           const double* __restrict__ w_old, const double* __restrict__ scale,
                                                           const double* __restrict__ centre, int P, int64_t j_per_split,
                                                           int64_t n_new_pad, double* __restrict__ den_part, const double* __restrict__ gate) {
    if (gate && gate_says_dmma(gate)) return;
    extern __shared__ double sm[];
    double* as = sm;                        // P * 128
    double* bs = as + (size_t)P * 128;      // 32 * P
    double* lw = bs + (size_t)32 * P;       // 32
    const int tid = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * 128 + tid;
    for (int p = 0; p < P; p++) as[p * 128 + tid] = (i < n_new) ? (th_new[(int64_t)p * ld_new + i] - centre[p]) * scale[p] : 0.0;
    const int64_t j0 = (int64_t)blockIdx.y * j_per_split, j1 = min(n_old, j0 + j_per_split);
    double den = 0;
    for (int64_t jt = j0; jt < j1; jt += 32) {
        __syncthreads();
        for (int idx = tid; idx < 32 * P; idx += 128) {
            const int jj = idx % 32, p = idx / 32;
            const int64_t j = jt + jj;
            bs[jj * P + p] = (j < j1) ? (th_old[(int64_t)p * ld_old + j] - centre[p]) * scale[p] : 0.0;
        }
        if (tid < 32) { const int64_t j = jt + tid; lw[tid] = (j < j1) ? -log(w_old[j]) : 1e300; }
        __syncthreads();
        for (int jj = 0; jj < 32; jj += 4) {
            double q0 = lw[jj], q1 = lw[jj + 1], q2 = lw[jj + 2], q3 = lw[jj + 3];
            for (int p = 0; p < P; p++) {
                const double a = as[p * 128 + tid];
                const double d0 = a - bs[(jj + 0) * P + p], d1 = a - bs[(jj + 1) * P + p];
                const double d2 = a - bs[(jj + 2) * P + p], d3 = a - bs[(jj + 3) * P + p];
                q0 = fma(d0, d0, q0); q1 = fma(d1, d1, q1); q2 = fma(d2, d2, q2); q3 = fma(d3, d3, q3);
            }
            den += exp_neg(q0) + exp_neg(q1) + exp_neg(q2) + exp_neg(q3);
        }
    }
    if (i < n_new) den_part[(int64_t)blockIdx.y * n_new_pad + i] = den;
}

// weight_i = numer_i / (C den_i) (den summed over j splits in fixed order); per-CTA partial sums of squares
__global__ void __launch_bounds__(256) weights_finalize_kernel(const double* __restrict__ den_part, int nsplit, const double* __restrict__ den_alt,
                                                               int nsplit_alt, const double* __restrict__ gate, int64_t n_new_pad,
                                                               int64_t n_new, const double* __restrict__ numer,
                                                               const double* __restrict__ cinv, const int* __restrict__ nanflag,
                                                               const int* __restrict__ poison, double* __restrict__ w_out,
                                                               double* __restrict__ ss_part) {
    __shared__ double red[32];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double w = 0.0;
    if (gate && !gate_says_dmma(gate)) { den_part = den_alt; nsplit = nsplit_alt; }     // auto mode: the pairwise kernel ran
    if (i < n_new) {
        double den = 0;
        for (int s = 0; s < nsplit; s++) den += den_part[(int64_t)s * n_new_pad + i];
        const double num = numer ? numer[i] : 1.0;
        w = num / (den / *cinv);
        if (*poison || nanflag[i]) w = __longlong_as_double(0x7ff8000000000000ll);
        w_out[i] = w;
    }
    const double ss = block_sum(w * w, red);
    if (threadIdx.x == 0) ss_part[blockIdx.x] = ss;
}

__global__ void sum_partials_kernel(const double* __restrict__ part, int n, double* __restrict__ out) {
    __shared__ double red[32];
    double s = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) *out = s;
}

// Eigen normalize(): if (squaredNorm > 0) v /= sqrt(squaredNorm)   (src/AbcUtil.cpp:583)
__global__ void scale_weights_kernel(double* __restrict__ w, int64_t n, const double* __restrict__ sumsq) {
    const double z = *sumsq;
    if (!(z > 0.0)) return;
    const double nrm = sqrt(z);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) w[i] /= nrm;
}

__global__ void fill_kernel(double* __restrict__ p, int64_t n, double v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

struct WPlan { int KS, NI, rows_per_cta; int64_t n_new_pad, n_old_pad; int row_blocks, nsplit, stages_per_split; bool dmma_ok; };

WPlan weights_plan(const abcb200_ctx* ctx, int64_t n_new, int64_t n_old, int P) {
    WPlan pl;
    const int ks = (P + 2 + 3) / 4;
    pl.dmma_ok = ks <= 16;
    pl.KS = ks <= 4 ? 4 : ks <= 6 ? 6 : ks <= 8 ? 8 : ks <= 12 ? 12 : 16;
    pl.NI = pl.KS <= 8 ? 4 : 2;
    pl.rows_per_cta = (W_THREADS / 32) * pl.NI * 8;
    pl.n_new_pad = (n_new + pl.rows_per_cta - 1) / pl.rows_per_cta * pl.rows_per_cta;
    pl.n_old_pad = (n_old + W_TJ - 1) / W_TJ * W_TJ;
    pl.row_blocks = (int)(pl.n_new_pad / pl.rows_per_cta);
    const int64_t total_stages = pl.n_old_pad / W_TJ;
    // Split the old particles so that (row blocks) x (splits) equal-sized tiles fill whole waves of the resident CTA slots
    // (two 256-thread CTAs per SM at 127 registers): 489 row blocks on 296 slots is 1.65 waves of work in 2 waves of time
    // (0.83, the 8-GPU share of the 1M x 1M case); 3 splits make 1467 tiles = 4.96 waves in 5 (0.99). The smallest split count
    // within 1 % of the best modelled efficiency is taken (a tile costs its stages + ~2 stages of ring fill).
    const int64_t slots = 2 * (int64_t)ctx->sm_count;
    int64_t smax = total_stages;
    if (smax > 128) smax = 128;
    if (smax < 1) smax = 1;
    int64_t want = 1;
    double best = -1.0;
    for (int64_t s = 1; s <= smax; s++) {
        const int64_t sps = (total_stages + s - 1) / s;
        const int64_t ns = (total_stages + sps - 1) / sps;            // the split count this request really gives
        const int64_t tiles = (int64_t)pl.row_blocks * ns;
        const int64_t waves = (tiles + slots - 1) / slots;
        // time ~ waves x (stages of the longest split + ring fill); work = row_blocks x total_stages
        const double eff = (double)pl.row_blocks * (double)total_stages / ((double)waves * (double)slots * (double)(sps + 2));
        if (eff > best * 1.01) { best = eff; want = s; }
    }
    pl.stages_per_split = (int)((total_stages + want - 1) / want);
    pl.nsplit = (int)((total_stages + pl.stages_per_split - 1) / pl.stages_per_split);
    return pl;
}

void diff_plan(const abcb200_ctx* ctx, int64_t n_new, int64_t n_old, int* row_blocks, int64_t* jps, int* nsplit) {
    *row_blocks = (int)((n_new + 127) / 128);
    int64_t want = (2 * (int64_t)ctx->sm_count + *row_blocks - 1) / *row_blocks;
    if (want < 1) want = 1;
    int64_t j = (n_old + want - 1) / want;
    j = (j + 31) / 32 * 32;
    *jps = j;
    *nsplit = (int)((n_old + j - 1) / j);
}

template <int KS, int NI>
int launch_dmma(abcb200_ctx* ctx, const WPlan& pl, const double* Apk, const double* Bpk, double* den_part, const double* gate) {
    const size_t smem = (size_t)W_STAGES * (W_TJ / 8) * KS * 32 * 8 + W_STAGES * 8 + 64;
    CUDA_TRY(ctx, cudaFuncSetAttribute(weights_dmma_kernel<KS, NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAUNCH(ctx, (weights_dmma_kernel<KS, NI>), dim3(pl.row_blocks, pl.nsplit), W_THREADS, smem, Apk, Bpk, pl.n_new_pad, pl.n_old_pad,
           pl.stages_per_split, den_part, gate);
    return ABCB200_OK;
}

}  // namespace

size_t weights_ws_bytes(const abcb200_ctx* ctx, int64_t n_new, int64_t n_old, int P) {
    const WPlan pl = weights_plan(ctx, n_new, n_old, P);
    size_t b = 0;
    b += align_up((size_t)pl.n_new_pad * pl.KS * 4 * 8, 256);
    b += align_up((size_t)pl.n_old_pad * pl.KS * 4 * 8, 256);
    int rb, ns; int64_t jps;
    diff_plan(ctx, n_new, n_old, &rb, &jps, &ns);
    b += align_up((size_t)pl.nsplit * pl.n_new_pad * 8, 256) + align_up((size_t)ns * pl.n_new_pad * 8, 256);   // auto mode keeps both
    b += align_up((size_t)pl.n_new_pad * 4, 256);
    b += 6 * align_up((size_t)P * 8, 256);
    b += align_up((size_t)((n_new + 255) / 256) * 8, 256);
    return b + 8192;
}

// Phase 1 of the update: derived constants, both operands packed into DMMA fragment order, and the conditioning maxima
// (job->scal[1] = max |a|^2 over THESE rows, job->scal[2] = max |b|^2) the device-side gate reads. A row-sharded caller
// (sharded.cu) max-reduces scal[1] over the ranks between the phases, so that every rank picks the same kernel.
int weights_pack(abcb200_ctx* ctx, const double* th_new, int64_t ld_new, int64_t n_new, const double* th_old, int64_t ld_old,
                 int64_t n_old, const double* w_old, const double* dv_old, int P, int algo, WeightsJob* job) {
    if (n_new < 1 || n_old < 1 || P < 1) ABC_FAIL(ctx, ABCB200_EINVAL, "weights: bad shape n_new=%lld n_old=%lld P=%d", (long long)n_new, (long long)n_old, P);
    if (ld_new < n_new || ld_old < n_old) ABC_FAIL(ctx, ABCB200_EINVAL, "weights: leading dimension smaller than the row count");
    const WPlan pl = weights_plan(ctx, n_new, n_old, P);
    WeightsJob& j = *job;
    j.th_new = th_new; j.ld_new = ld_new; j.n_new = n_new; j.th_old = th_old; j.ld_old = ld_old; j.n_old = n_old; j.w_old = w_old;
    j.P = P; j.algo = algo;
    j.scale = ws_new<double>(ctx, P);
    j.centre = ws_new<double>(ctx, P);
    double* colmin = ws_new<double>(ctx, P);
    double* colmax = ws_new<double>(ctx, P);
    j.scal = ws_new<double>(ctx, 4);      // [0] 1/C, [1] max |a|^2, [2] max |b|^2
    j.poison = ws_new<int>(ctx, 1);
    j.nanflag = ws_new<int>(ctx, pl.n_new_pad);
    j.Apk = ws_new<double>(ctx, (size_t)pl.n_new_pad * pl.KS * 4);
    j.Bpk = ws_new<double>(ctx, (size_t)pl.n_old_pad * pl.KS * 4);
    j.nfin = (int)((n_new + 255) / 256);
    j.ss_part = ws_new<double>(ctx, j.nfin);
    if (!j.scale || !j.centre || !colmin || !colmax || !j.scal || !j.poison || !j.nanflag || !j.Apk || !j.Bpk || !j.ss_part)
        ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in weights");
    CUDA_TRY(ctx, cudaMemsetAsync(j.scal, 0, 4 * sizeof(double), ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(j.poison, 0, sizeof(int), ctx->stream));
    LAUNCH(ctx, weights_prep_kernel, P, 256, 0, th_old, ld_old, n_old, dv_old, P, j.scale, j.centre, j.scal, colmin, colmax, j.poison);
    LAUNCH(ctx, pack_old_kernel, (unsigned)((pl.n_old_pad + 127) / 128), 128, 0, th_old, ld_old, n_old, pl.n_old_pad, w_old, j.scale, j.centre, P,
           pl.KS, j.Bpk, j.scal + 2, j.poison);
    LAUNCH(ctx, pack_new_kernel, (unsigned)((pl.n_new_pad + 127) / 128), 128, 0, th_new, ld_new, n_new, pl.n_new_pad, j.scale, j.centre, dv_old,
           colmin, colmax, P, pl.KS, j.Apk, j.scal + 1, j.nanflag);
    return ABCB200_OK;
}

// Phase 2: the pair kernel(s), then weight_i = numer_i / den_i and the sum of squares of these rows.
int weights_eval(abcb200_ctx* ctx, const WeightsJob* job, const double* numer, double* w_out, double* sumsq_out) {
    const WeightsJob& j = *job;
    const WPlan pl = weights_plan(ctx, j.n_new, j.n_old, j.P);
    const int P = j.P;
    // algo 1: pairwise-difference kernel (the reference's formulation); algo 2: DMMA inner-product kernel; algo 0 (auto): both
    // are enqueued and the conditioning of the expanded exponent, measured by the pack kernels, decides on the device.
    const bool want_dmma = pl.dmma_ok && j.algo != 1, want_diff = !pl.dmma_ok || j.algo != 2;
    const double* gate = (want_dmma && want_diff) ? j.scal : nullptr;
    double *den_dmma = nullptr, *den_diff = nullptr;
    int nsplit_dmma = 0, nsplit_diff = 0;
    kernel_begin(ctx, 7);
    if (want_dmma) {
        nsplit_dmma = pl.nsplit;
        den_dmma = ws_new<double>(ctx, (size_t)nsplit_dmma * pl.n_new_pad);
        if (!den_dmma) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in weights");
        switch (pl.KS) {
            case 4: ABC_TRY((launch_dmma<4, 4>(ctx, pl, j.Apk, j.Bpk, den_dmma, gate))); break;
            case 6: ABC_TRY((launch_dmma<6, 4>(ctx, pl, j.Apk, j.Bpk, den_dmma, gate))); break;
            case 8: ABC_TRY((launch_dmma<8, 4>(ctx, pl, j.Apk, j.Bpk, den_dmma, gate))); break;
            case 12: ABC_TRY((launch_dmma<12, 2>(ctx, pl, j.Apk, j.Bpk, den_dmma, gate))); break;
            default: ABC_TRY((launch_dmma<16, 2>(ctx, pl, j.Apk, j.Bpk, den_dmma, gate))); break;
        }
    }
    if (want_diff) {
        int row_blocks; int64_t jps;
        diff_plan(ctx, j.n_new, j.n_old, &row_blocks, &jps, &nsplit_diff);
        den_diff = ws_new<double>(ctx, (size_t)nsplit_diff * pl.n_new_pad);
        if (!den_diff) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in weights");
        const size_t smem = sizeof(double) * ((size_t)P * 128 + 32 * (size_t)P + 32);
        if (smem > (size_t)ctx->smem_optin) ABC_FAIL(ctx, ABCB200_EINVAL, "weights: P=%d too large for the pairwise kernel", P);
        CUDA_TRY(ctx, cudaFuncSetAttribute(weights_diff_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LAUNCH(ctx, weights_diff_kernel, dim3(row_blocks, nsplit_diff), 128, smem, j.th_new, j.ld_new, j.n_new, j.th_old, j.ld_old, j.n_old, j.w_old, j.scale,
               j.centre, P, jps, pl.n_new_pad, den_diff, gate);
    }
    const double* den_part = want_dmma ? den_dmma : den_diff;
    const int nsplit = want_dmma ? nsplit_dmma : nsplit_diff;
    kernel_end(ctx, 7);
    LAUNCH(ctx, weights_finalize_kernel, j.nfin, 256, 0, den_part, nsplit, den_diff, nsplit_diff, gate, pl.n_new_pad, j.n_new, numer, j.scal, j.nanflag, j.poison,
           w_out, j.ss_part);
    LAUNCH(ctx, sum_partials_kernel, 1, 256, 0, j.ss_part, j.nfin, sumsq_out);
    return ABCB200_OK;
}

int weights_unnorm_dev(abcb200_ctx* ctx, const double* numer, const double* th_new, int64_t ld_new, int64_t n_new,
                       const double* th_old, int64_t ld_old, int64_t n_old, const double* w_old, const double* dv_old, int P,
                       int algo, double* w_out, double* sumsq_out) {
    if (n_new < 0) ABC_FAIL(ctx, ABCB200_EINVAL, "weights: bad shape n_new=%lld", (long long)n_new);
    if (n_new == 0) { CUDA_TRY(ctx, cudaMemsetAsync(sumsq_out, 0, 8, ctx->stream)); return ABCB200_OK; }
    WeightsJob job;
    ABC_TRY(weights_pack(ctx, th_new, ld_new, n_new, th_old, ld_old, n_old, w_old, dv_old, P, algo, &job));
    return weights_eval(ctx, &job, numer, w_out, sumsq_out);
}

int launch_scale_weights(abcb200_ctx* ctx, double* w, int64_t n, const double* sumsq) {
    const int grid = (int)max((int64_t)1, min((n + 255) / 256, (int64_t)(4 * ctx->sm_count)));
    LAUNCH(ctx, scale_weights_kernel, grid, 256, 0, w, n, sumsq);
    return ABCB200_OK;
}

int launch_fill(abcb200_ctx* ctx, double* p, int64_t n, double v) {
    const int grid = (int)max((int64_t)1, min((n + 255) / 256, (int64_t)(4 * ctx->sm_count)));
    LAUNCH(ctx, fill_kernel, grid, 256, 0, p, n, v);
    return ABCB200_OK;
}
