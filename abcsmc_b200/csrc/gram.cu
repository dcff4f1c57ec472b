// gram.cu — XX = X^T X and XY = X^T Y in ONE pass over the training rows (SURVEY.md §8 row a4 / S2').
//
// Reference: PLS::Model::plsr forms XY = X^T Y (lib/PLS/src/pls.cpp:396) and, for KERNEL_TYPE2, XX = X^T X (:398).
// pls_gram.cu needs both, so they are one FP64 tensor-core contraction over the N_tr rows:
//
//   * 8x8 output tiles; the tile columns are the X tiles [0, nTx) followed by the Y tiles [nTx, nT). Only the pairs
//     (ta, tb) with ta < nTx and tb >= ta are computed (upper triangle of XX, all of XY); XX is mirrored on output,
//     which also makes it bitwise symmetric.
//   * A CTA (16 warps) owns a block of tile rows [a0, a1) and a chunk of data rows. The needed columns of a stage
//     of RT data rows are brought into a shared-memory ring by cp.async.bulk (one 8*RT-byte column segment per copy,
//     mbarrier complete_tx; SASS UBLKCP), issued by the thread that owns the column one stage after the slot was
//     released, so no warp ever waits for a copy it issues. Columns are padded to RT + 4 doubles: the DMMA fragment
//     loads (lane = 4 columns x 4 rows per half warp) are then bank-conflict free.
//   * A warp keeps a contiguous run of <= MAXT tile pairs in registers for the CTA's lifetime (DMMA.8x8x4
//     accumulators), reloading the A fragment only when the tile row changes. When a row block has few pairs
//     (small K) the 16 warps form G groups that take alternate 4-row steps of every stage and are summed at the end in
//     a fixed order. X is read from HBM once (row blocks of one chunk run side by side and share it through L2).
//   * Per-chunk partial tiles are reduced in a fixed order by gram_reduce_kernel: deterministic results.
//
// Algorithmic work: 64 * 2 * 4 flop per tile pair and 4 rows -> n * (nTx (nTx + 1) / 2 + nTx nTy) * 128 flop
// (about n K (K + 1) + 2 n K M); 8 n (K + M) bytes.
#include <stdlib.h>

#include <algorithm>

#include "kernels.cuh"

namespace {

constexpr int GR_T = 512;          // threads
constexpr int GR_W = GR_T / 32;    // warps
constexpr int GR_NS = 4;           // ring stages
constexpr int GR_MAXRB = 128;      // tile blocks per launch
constexpr int GR_MAXSLOT = 72;     // ring slots (8-column tiles) a block may stage

// A block = tile rows [a0, a1) x tile columns [b0, b1). Diagonal kind (b0 == a0): the pairs (ta, tb), ta <= tb < b1 - the triangle of
// the row block plus the rectangle to its right up to b1. Off-diagonal kind (b0 >= a1): the full rectangle. A CTA stages only the
// columns of ITS tiles: [a0, b1) for a diagonal block, [a0, a1) followed by [b0, b1) otherwise. (Round 1 gave every row block all
// the columns to its right: at K = 500 each stage was 560 bulk copies for 280 tile pairs and X was read 2.7 times from HBM.)
struct GramTileArgs {
    const double* X; const double* Y;
    int64_t ldx, ldy, n, rows_per_chunk;
    int K, M, nTx, nT, n_rb, G, TPW;
    int NS;                        // ring stages in use (2 .. GR_NS): wide operands trade depth for longer column segments per copy
    unsigned short a0[GR_MAXRB], a1[GR_MAXRB], b0[GR_MAXRB], b1[GR_MAXRB];
    int pair_base[GR_MAXRB + 1];   // prefix sums of pairs per block
    double* partial;               // [chunk][pair][64]
    int total_pairs;
};

// pair number `idx` of a block (row-major over its tile rows) -> (ta, tb)
__host__ __device__ __forceinline__ void gram_pair(int a0, int a1, int b0, int b1, int idx, int& ta, int& tb) {
    const bool diag = b0 == a0;
    ta = a0;
    while (ta < a1 - 1 && idx >= b1 - (diag ? ta : b0)) { idx -= b1 - (diag ? ta : b0); ta++; }
    tb = (diag ? ta : b0) + idx;
}

constexpr int GR_BT = 6;           // B fragments in flight per batch
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int MAXT, int RT>
__global__ void __launch_bounds__(GR_T, 1) gram_kernel(const GramTileArgs p) {
    constexpr int LD = RT + 4;             // padded column length (doubles)
    constexpr int TILE_D = 8 * LD;         // doubles per tile column group
    constexpr int KS = RT / 4;
    extern __shared__ __align__(128) double sm[];
    __shared__ __align__(8) uint64_t full[GR_NS], empty[GR_NS];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int rb = blockIdx.x;
    const int NS = p.NS;
    const int a0 = p.a0[rb], a1 = p.a1[rb], b0 = p.b0[rb], b1 = p.b1[rb];
    const bool diag = b0 == a0;
    const int nra = a1 - a0;                                   // ring slots of the row tiles (an off-diagonal block's column tiles follow)
    const int nTx = p.nTx;
    const int nslot = diag ? b1 - a0 : nra + (b1 - b0);
    const int ncol_s = 8 * nslot;
    const int stage_d = ncol_s * LD;
    const int Tb = p.pair_base[rb + 1] - p.pair_base[rb];
    const int G = p.G, Wg = GR_W / G, TPW = p.TPW;
    const int gi = wid / Wg, wi = wid - gi * Wg;
    const int64_t r0 = (int64_t)blockIdx.y * p.rows_per_chunk;
    const int64_t r1 = min(p.n, r0 + p.rows_per_chunk);
    const int nst = (r1 > r0) ? (int)((r1 - r0 + RT - 1) / RT) : 0;

    // ---- this warp's run of tile pairs, packed as ring slots (slot of ta) | (slot of tb) << 8 ------------------------
    uint32_t pk[MAXT / 2];          // two pairs per register: bytes (ia, ib) of pair 2u, then of pair 2u + 1
    int cnt;
    {
        const int first = wi * TPW;
        cnt = max(0, min(TPW, Tb - first));
        int ta, tb;
        gram_pair(a0, a1, b0, b1, min(first, Tb - 1), ta, tb);   // idle warp (cnt == 0): any valid pair, its slots are never stored
#pragma unroll
        for (int t = 0; t < MAXT; t++) {
            const uint32_t v = (uint32_t)(ta - a0) | ((uint32_t)(diag ? tb - a0 : nra + tb - b0) << 8);
            if (t & 1) pk[t / 2] |= v << 16; else pk[t / 2] = v;
            if (t + 1 < cnt) { tb++; if (tb == b1) { ta++; tb = diag ? ta : b0; } }
        }
    }

    // ---- the columns this thread feeds (column c of the ring: tile a0 + c/8, lane c%8) -----------------------------
    const int n_issue_w = min(GR_W, (ncol_s + 31) / 32);
    const double* src[2];
    uint32_t wbytes = 0;
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
        const int c = tid + rr * GR_T;
        const double* s = nullptr;
        if (c < ncol_s) {
            const int slot = c >> 3, j = c & 7;
            const int t = (diag || slot < nra) ? a0 + slot : b0 + (slot - nra);
            if (t < nTx) { const int col = 8 * t + j; if (col < p.K) s = p.X + (int64_t)col * p.ldx; }
            else { const int col = 8 * (t - nTx) + j; if (col < p.M) s = p.Y + (int64_t)col * p.ldy; }
        }
        src[rr] = s;
        wbytes += (uint32_t)__popc(__ballot_sync(0xffffffffu, s != nullptr)) * (RT * 8);
    }
    if (tid == 0) {
        for (int s = 0; s < NS; s++) { mbar_init(&full[s], n_issue_w); mbar_init(&empty[s], GR_W); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // columns without a source (padding of the last X / Y tile) stay zero in every slot
    for (int rr = 0; rr < 2; rr++) {
        const int c = tid + rr * GR_T;
        if (c < ncol_s && src[rr] == nullptr)
            for (int s = 0; s < NS; s++)
                for (int r = 0; r < RT; r++) sm[(size_t)s * stage_d + c * LD + r] = 0.0;
    }
    __syncthreads();

    auto issue = [&](int st) {
        if (wid >= n_issue_w) return;
        const int slot = st % NS;
        if (st >= NS) mbar_wait(&empty[slot], (uint32_t)((st / NS - 1) & 1));
        const int64_t row = r0 + (int64_t)st * RT;
        double* dst = sm + (size_t)slot * stage_d;
        if (row + RT <= r1) {
            if (lane == 0) mbar_expect_tx(&full[slot], wbytes);
            __syncwarp();
#pragma unroll
            for (int rr = 0; rr < 2; rr++)
                if (src[rr]) tma_bulk_g2s(dst + (tid + rr * GR_T) * LD, src[rr] + row, RT * 8, &full[slot]);
        } else {   // ragged last stage: plain loads, rows beyond r1 contribute zero
#pragma unroll
            for (int rr = 0; rr < 2; rr++)
                if (src[rr]) {
                    double* d = dst + (tid + rr * GR_T) * LD;
                    for (int r = 0; r < RT; r++) d[r] = (row + r < r1) ? src[rr][row + r] : 0.0;
                }
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[slot]);
        }
    };

    const uint32_t ring32 = smem_u32(sm);
    double acc[MAXT][2];
#pragma unroll
    for (int t = 0; t < MAXT; t++) acc[t][0] = acc[t][1] = 0.0;

    for (int st = 0; st < min(NS - 1, nst); st++) issue(st);
    for (int st = 0; st < nst; st++) {
        if (st + NS - 1 < nst) issue(st + NS - 1);
        const int slot = st % NS;
        mbar_wait(&full[slot], (uint32_t)((st / NS) & 1));
        const uint32_t sp = ring32 + (uint32_t)slot * (uint32_t)(stage_d * 8) + (uint32_t)((g * LD + q) * 8);
#pragma unroll 1
        for (int ks = gi; ks < KS; ks += G) {
            const uint32_t s4 = sp + 32u * (uint32_t)ks;
            double a = 0.0;
            // every slot of the run is executed (slots beyond cnt repeat the last pair; their accumulators are never stored):
            // no branches between the DMMAs, the B fragments of a batch are in flight together
#pragma unroll
            for (int t0 = 0; t0 < MAXT; t0 += GR_BT) {
                double b[GR_BT];
#pragma unroll
                for (int u = 0; u < GR_BT; u++) {
                    const int t = t0 + u;
                    if (t < MAXT) b[u] = lds_f64(s4 + ((pk[t / 2] >> (16 * (t & 1) + 8)) & 0xffu) * (uint32_t)(TILE_D * 8));
                }
#pragma unroll
                for (int u = 0; u < GR_BT; u++) {
                    const int t = t0 + u;
                    if (t < MAXT) {
                        const uint32_t ia = (pk[t / 2] >> (16 * (t & 1))) & 0xffu;
                        const int tp = t - (t > 0);
                        if (t == 0 || ia != ((pk[tp / 2] >> (16 * (tp & 1))) & 0xffu)) a = lds_f64(s4 + ia * (uint32_t)(TILE_D * 8));
                        dmma884(acc[t][0], acc[t][1], a, b[u]);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
    }

    // ---- cross-group sum (fixed order) and the chunk's partial tiles -------------------------------------------
    double* out = p.partial + ((size_t)blockIdx.y * p.total_pairs + p.pair_base[rb]) * 64;
    if (G == 1) {
#pragma unroll
        for (int t = 0; t < MAXT; t++)
            if (t < cnt) *(double2*)(out + (size_t)(wi * TPW + t) * 64 + 2 * lane) = make_double2(acc[t][0], acc[t][1]);
    } else {
        __syncthreads();                       // every stage consumed: the ring is free
        const int per_g = Wg * MAXT * 64;      // doubles per group
#pragma unroll
        for (int t = 0; t < MAXT; t++)
            *(double2*)(sm + (size_t)gi * per_g + (wi * MAXT + t) * 64 + 2 * lane) = make_double2(acc[t][0], acc[t][1]);
        __syncthreads();
        for (int e = tid; e < per_g; e += GR_T) {
            const int w = e / (MAXT * 64), t = (e / 64) % MAXT;
            if (w * TPW + t < Tb && t < TPW) {
                double s = 0.0;
                for (int gg = 0; gg < G; gg++) s += sm[(size_t)gg * per_g + e];
                out[(size_t)(w * TPW + t) * 64 + (e & 63)] = s;
            }
        }
    }
}

// XX / XY from the per-chunk partial tiles, summed in a fixed order. One CTA per tile pair: 64 elements x GR_RS chunk slices
// (a slice = a contiguous run of chunks; short latency chain when there are ~148 chunks), slices combined pairwise.
constexpr int GR_RS = 4;
__global__ void __launch_bounds__(64 * GR_RS) gram_reduce_kernel(const GramTileArgs p, int nchunk, double* __restrict__ XX, double* __restrict__ XY) {
    __shared__ double part[GR_RS][64];
    const int pair = blockIdx.x, e = threadIdx.x & 63, slice = threadIdx.x >> 6;
    int rb = 0;
    while (pair >= p.pair_base[rb + 1]) rb++;
    int ta, tb;
    gram_pair(p.a0[rb], p.a1[rb], p.b0[rb], p.b1[rb], pair - p.pair_base[rb], ta, tb);
    const double* src = p.partial + (size_t)pair * 64 + e;
    const size_t stride = (size_t)p.total_pairs * 64;
    const int per = (nchunk + GR_RS - 1) / GR_RS, c_end = min(nchunk, (slice + 1) * per);
    double a[4] = {0, 0, 0, 0};
    int c = slice * per;
    for (; c + 4 <= c_end; c += 4) {
#pragma unroll
        for (int u = 0; u < 4; u++) a[u] += src[(size_t)(c + u) * stride];
    }
    for (int u = 0; c + u < c_end; u++) a[u] += src[(size_t)(c + u) * stride];
    part[slice][e] = (a[0] + a[1]) + (a[2] + a[3]);
    __syncthreads();
    if (slice != 0) return;
    const double v = (part[0][e] + part[1][e]) + (part[2][e] + part[3][e]);
    // fragment order: element e = 2 * lane + j -> row g = lane >> 2, column 2 * (lane & 3) + j
    const int l = e >> 1, j = e & 1;
    const int row = 8 * ta + (l >> 2);
    if (row >= p.K) return;
    if (tb < p.nTx) {
        const int col = 8 * tb + 2 * (l & 3) + j;
        if (col >= p.K) return;
        if (ta == tb) { if (col >= row) { XX[(size_t)col * p.K + row] = v; XX[(size_t)row * p.K + col] = v; } }
        else { XX[(size_t)col * p.K + row] = v; XX[(size_t)row * p.K + col] = v; }
    } else {
        const int col = 8 * (tb - p.nTx) + 2 * (l & 3) + j;
        if (col < p.M) XY[(size_t)col * p.K + row] = v;
    }
}

struct GramPlan {
    GramTileArgs a;
    int nchunk, maxt, rt;
    size_t smem;
};

int gram_plan(const abcb200_ctx* ctx, int64_t n, int K, int M, GramPlan* pl) {
    GramTileArgs& a = pl->a;
    a.K = K; a.M = M; a.n = n;
    a.nTx = (K + 7) / 8; a.nT = a.nTx + (M + 7) / 8;
    // Blocks of at most 16 warps x 20 tile pairs. Everything in one diagonal block when it fits (K <= ~150); otherwise tile-row blocks
    // of height h, each cut into a diagonal block (its triangle + as many columns to the right as fill `cap` pairs) and off-diagonal
    // blocks that share the remaining columns evenly: a CTA then stages 8 (h + w) columns for h w pairs instead of every column to
    // its right. All CTAs of the launch run side by side (blocks x row chunks <= SMs), so the launch takes what its largest block
    // takes: (h, cap) is the pair that minimises (pairs of the largest block) x (rows per chunk).
    const int MAXP = GR_W * 20;
    int nb = 0, pairs = 0, maxTb = 0, max_slots = 0;
    auto add_block = [&](int a0, int a1, int b0, int b1) -> bool {
        if (nb >= GR_MAXRB) return false;
        int np = 0;
        for (int ta = a0; ta < a1; ta++) np += b1 - ((b0 == a0) ? ta : b0);
        if (np <= 0 || np > MAXP) return false;
        a.a0[nb] = (unsigned short)a0; a.a1[nb] = (unsigned short)a1; a.b0[nb] = (unsigned short)b0; a.b1[nb] = (unsigned short)b1;
        pairs += np; a.pair_base[++nb] = pairs;
        if (np > maxTb) maxTb = np;
        const int slots = (b0 == a0) ? b1 - a0 : (a1 - a0) + (b1 - b0);
        if (slots > max_slots) max_slots = slots;
        return true;
    };
    auto build = [&](int h, int cap) -> bool {
        nb = 0; pairs = 0; maxTb = 0; max_slots = 0;
        a.pair_base[0] = 0;
        for (int a0 = 0; a0 < a.nTx; a0 += h) {
            const int a1 = a0 + h < a.nTx ? a0 + h : a.nTx, hh = a1 - a0;
            const int tri = hh * (hh + 1) / 2;
            if (tri > cap) return false;
            int d1 = a1 + (cap - tri) / hh;
            if (d1 > a0 + GR_MAXSLOT) d1 = a0 + GR_MAXSLOT;            // ring width (and the byte-packed slot numbers) stay bounded
            if (d1 > a.nT) d1 = a.nT;
            if (!add_block(a0, a1, a0, d1)) return false;
            int w = cap / hh;
            if (w > GR_MAXSLOT - hh) w = GR_MAXSLOT - hh;
            const int rest = a.nT - d1;
            if (rest > 0) {
                const int nsplit = (rest + w - 1) / w, we = (rest + nsplit - 1) / nsplit;       // even widths
                for (int b0 = d1; b0 < a.nT; b0 += we)
                    if (!add_block(a0, a1, b0, b0 + we < a.nT ? b0 + we : a.nT)) return false;
            }
        }
        return true;
    };
    const int all_pairs = a.nTx * (a.nTx + 1) / 2 + a.nTx * (a.nT - a.nTx);
    if (all_pairs <= MAXP) {
        a.pair_base[0] = 0;
        if (!add_block(0, a.nTx, 0, a.nT)) return -1;
    } else {
        int best_h = 0, best_cap = 0;
        double best_cost = 1e300;
        for (int h = 8; h <= 20; h += 4)
            for (int cap = MAXP; cap >= 96; cap -= 16) {
                if (!build(h, cap)) continue;
                const int nch = ctx->sm_count / nb > 0 ? ctx->sm_count / nb : 1;
                const double cost = (double)maxTb * (double)((n + nch - 1) / nch) * (1.0 + 0.15 * (double)max_slots / (double)maxTb * 8.0);   // + the staging a pair pays for
                if (cost < best_cost) { best_cost = cost; best_h = h; best_cap = cap; }
            }
        if (best_h == 0 || !build(best_h, best_cap)) return -1;
    }
    const int rb = nb;
    a.n_rb = nb; a.total_pairs = pairs;
    // warps per group: as few as keep a warp's run within 20 pairs, but at least 6 pairs per warp when there is a choice
    int Wg = GR_W;
    while (Wg > 1 && (maxTb + Wg / 2 - 1) / (Wg / 2) <= 12) Wg /= 2;
    a.G = GR_W / Wg;
    a.TPW = (maxTb + Wg - 1) / Wg;
    pl->maxt = (a.TPW + 1) / 2 * 2;
    const size_t ncol0 = (size_t)8 * max_slots;       // widest ring of any block
    const size_t budget = (size_t)ctx->smem_optin - 256;
    const size_t red = (a.G > 1) ? (size_t)GR_W * pl->maxt * 64 * 8 : 0;
    int rt = a.G * 4 < 16 ? 16 : a.G * 4;
    // Every column of a stage is one cp.async.bulk of 8 * rt bytes, and small bulk copies are issue bound (K = 500: 560 copies of
    // 64 bytes per 8 rows left the DMMA pipe at 22 %). Wide operands therefore keep rt = 16 and give up ring depth (4 -> 3 -> 2
    // stages) before they fall back to 8-row stages.
    if (rt == 16 && a.G == 1 && (size_t)GR_NS * ncol0 * (32 + 4) * 8 <= budget) rt = 32;     // longer segments when four 32-row stages fit
    int ns = GR_NS;
    while (rt == 16 && ns > 2 && (size_t)ns * ncol0 * (16 + 4) * 8 > budget) ns--;
    if (rt == 16 && (size_t)ns * ncol0 * (16 + 4) * 8 > budget && a.G <= 2) { rt = 8; ns = GR_NS; }
    pl->rt = rt;
    a.NS = ns;
    size_t ring = (size_t)ns * ncol0 * (size_t)(rt + 4) * 8;
    pl->smem = ring > red ? ring : red;
    if (pl->smem > budget) return -1;
    // chunks: one CTA per SM
    int nchunk = ctx->sm_count / rb;
    if (nchunk < 1) nchunk = 1;
    int64_t rpc = (n + nchunk - 1) / nchunk;
    rpc = (rpc + rt - 1) / rt * rt;
    if (rpc < 4 * rt) rpc = 4 * rt;
    a.rows_per_chunk = rpc;
    pl->nchunk = (int)((n + rpc - 1) / rpc);
    return 0;
}

template <int MAXT>
cudaError_t gram_launch_rt(const GramPlan& pl, cudaStream_t st) {
    const dim3 grid(pl.a.n_rb, pl.nchunk);
#define GR_CASE(RT_)                                                                                                    \
    case RT_: {                                                                                                         \
        cudaError_t e = cudaFuncSetAttribute(gram_kernel<MAXT, RT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem); \
        if (e != cudaSuccess) return e;                                                                                 \
        gram_kernel<MAXT, RT_><<<grid, GR_T, pl.smem, st>>>(pl.a);                                                      \
        return cudaGetLastError();                                                                                      \
    }
    switch (pl.rt) {
        GR_CASE(8)
        GR_CASE(16)
        GR_CASE(32)
        GR_CASE(64)
    }
#undef GR_CASE
    return cudaErrorInvalidValue;
}

}  // namespace

size_t gram_ws_bytes(const abcb200_ctx* ctx, int64_t n, int K, int M) {
    GramPlan pl;
    if (gram_plan(ctx, n, K, M, &pl) != 0) return 0;
    return align_up((size_t)pl.nchunk * pl.a.total_pairs * 64 * sizeof(double), 256) + 256;
}

// XX (K x K, ld K, both triangles) = X^T X and XY (K x M, ld K) = X^T Y over n rows. X and Y columns must be 16-byte
// aligned (even leading dimensions, even row offset): the device copies made by api.cu are.
int launch_gram(abcb200_ctx* ctx, const double* X, int64_t ldx, int K, const double* Y, int64_t ldy, int M, int64_t n, double* XX, double* XY) {
    GramPlan pl;
    if (gram_plan(ctx, n, K, M, &pl) != 0) ABC_FAIL(ctx, ABCB200_EINVAL, "gram: K=%d M=%d outside the tiled Gram kernel's range", K, M);
    if (((uintptr_t)X & 15) || ((uintptr_t)Y & 15) || (ldx & 1) || (ldy & 1)) ABC_FAIL(ctx, ABCB200_EINVAL, "gram: operands must have 16-byte aligned columns");
    pl.a.X = X; pl.a.Y = Y; pl.a.ldx = ldx; pl.a.ldy = ldy;
    pl.a.partial = ws_new<double>(ctx, (size_t)pl.nchunk * pl.a.total_pairs * 64);
    if (!pl.a.partial) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in launch_gram");
    if (getenv("ABCB200_DEBUG")) {
        int mx = 0;
        for (int b = 0; b < pl.a.n_rb; b++) mx = std::max(mx, pl.a.pair_base[b + 1] - pl.a.pair_base[b]);
        fprintf(stderr, "[abcb200] gram K=%d M=%d n=%lld: %d blocks (largest %d of %d pairs, first block rows [%d,%d) cols [%d,%d)), %d chunks of %lld rows, rt=%d ns=%d maxt=%d G=%d\n", K, M,
                (long long)n, pl.a.n_rb, mx, pl.a.total_pairs, (int)pl.a.a0[0], (int)pl.a.a1[0], (int)pl.a.b0[0], (int)pl.a.b1[0], pl.nchunk, (long long)pl.a.rows_per_chunk, pl.rt, pl.a.NS, pl.maxt, pl.a.G);
    }
    cudaError_t e;
    kernel_begin(ctx, 1);
    switch (pl.maxt) {
        case 2: e = gram_launch_rt<2>(pl, ctx->stream); break;
        case 4: e = gram_launch_rt<4>(pl, ctx->stream); break;
        case 6: e = gram_launch_rt<6>(pl, ctx->stream); break;
        case 8: e = gram_launch_rt<8>(pl, ctx->stream); break;
        case 10: e = gram_launch_rt<10>(pl, ctx->stream); break;
        case 12: e = gram_launch_rt<12>(pl, ctx->stream); break;
        case 14: e = gram_launch_rt<14>(pl, ctx->stream); break;
        case 16: e = gram_launch_rt<16>(pl, ctx->stream); break;
        case 18: e = gram_launch_rt<18>(pl, ctx->stream); break;
        default: e = gram_launch_rt<20>(pl, ctx->stream); break;
    }
    kernel_end(ctx, 1);
    ctx->launches++;
    CUDA_TRY(ctx, e);
    LAUNCH(ctx, gram_reduce_kernel, pl.a.total_pairs, 64 * GR_RS, 0, pl.a, pl.nchunk, XX, XY);
    return ABCB200_OK;
}
