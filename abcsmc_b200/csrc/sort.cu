// sort.cu — S6 ordering and the batched segmented sort behind the Wilcoxon tests (a8).
//
// Reference: PLS::ordered, lib/PLS/include/PLS/pls.h:58-69 — std::sort of indices by value with strict <
// (not stable: tie order is whatever libstdc++ introsort produces). This implementation is a STABLE LSD radix
// sort started from ascending indices, so ties come out in ascending particle index (documented deviation for
// exact ties only; continuous inputs have none).
//
// Device design: 8-bit digits, three kernels per pass (per-tile histogram, per-segment exclusive scan, stable
// scatter with warp match-any ranking). Segments are equal length and independent, so a batch of signed-rank
// tests is sorted by one set of launches with gridDim.y = number of segments. Bound: HBM/L2;
// algorithmic bytes per key per pass: 8 (histogram read) + 8 (scatter read) + 8 (scatter write) (+4+4 with payload).
#include "kernels.cuh"

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_KPT = 8;
constexpr int RS_TILE = RS_THREADS * RS_KPT;   // 2048 keys per CTA
constexpr int RS_WARPS = RS_THREADS / 32;

__global__ void __launch_bounds__(RS_THREADS) radix_hist_kernel(const uint64_t* __restrict__ keys, int64_t seg_len, int tiles,
                                                                int shift, uint32_t* __restrict__ hist,
                                                                const int* __restrict__ seg_valid) {
    const int seg = blockIdx.y, tile = blockIdx.x;
    if (seg_valid && !seg_valid[seg]) return;
    __shared__ uint32_t cnt[256];
    cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t* k = keys + (int64_t)seg * seg_len;
    const int64_t base = (int64_t)tile * RS_TILE;
#pragma unroll
    for (int j = 0; j < RS_KPT; j++) {
        const int64_t i = base + (int64_t)j * RS_THREADS + threadIdx.x;
        if (i < seg_len) atomicAdd(&cnt[(uint32_t)(k[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[((int64_t)seg * tiles + tile) * 256 + threadIdx.x] = cnt[threadIdx.x];   // [seg][tile][digit]: coalesced
}

// exclusive scan over the (digit-major, tile-minor) ordering of one segment's counts, in place; storage is
// [tile][digit] so that every access of the 256 threads (one per digit) is coalesced
__global__ void __launch_bounds__(256) radix_scan_kernel(uint32_t* __restrict__ hist, int tiles, const int* __restrict__ seg_valid) {
    const int seg = blockIdx.x;
    if (seg_valid && !seg_valid[seg]) return;
    __shared__ uint32_t wsum[8];
    uint32_t* row = hist + (int64_t)seg * tiles * 256 + threadIdx.x;
    uint32_t total = 0;
    for (int t = 0; t < tiles; t++) total += row[(int64_t)t * 256];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    uint32_t wbase = 0;
    for (int w = 0; w < wid; w++) wbase += wsum[w];
    uint32_t run = wbase + incl - total;
    for (int t = 0; t < tiles; t++) { const uint32_t c = row[(int64_t)t * 256]; row[(int64_t)t * 256] = run; run += c; }
}

template <bool HAS_VAL>
__global__ void __launch_bounds__(RS_THREADS) radix_scatter_kernel(const uint64_t* __restrict__ keys_in, uint64_t* __restrict__ keys_out,
                                                                   const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ vals_out,
                                                                   int64_t seg_len, int tiles, int shift,
                                                                   const uint32_t* __restrict__ offs, const int* __restrict__ seg_valid) {
    const int seg = blockIdx.y, tile = blockIdx.x;
    if (seg_valid && !seg_valid[seg]) return;
    __shared__ uint32_t cnt[RS_WARPS][256];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const uint64_t* kin = keys_in + (int64_t)seg * seg_len;
    const int64_t wbase = (int64_t)tile * RS_TILE + (int64_t)wid * (RS_KPT * 32);
    uint64_t key[RS_KPT];
    uint32_t val[RS_KPT];
    uint32_t lrank[RS_KPT];
#pragma unroll
    for (int r = 0; r < RS_KPT; r++) {
        const int64_t i = wbase + r * 32 + lane;
        const bool valid = i < seg_len;
        key[r] = valid ? kin[i] : 0ull;
        if (HAS_VAL) val[r] = valid ? vals_in[(int64_t)seg * seg_len + i] : 0u;
        const uint32_t active = __ballot_sync(0xffffffffu, valid);
        uint32_t below = 0, old = 0;
        int leader = lane;
        if (valid) {
            const uint32_t d = (uint32_t)(key[r] >> shift) & 255u;
            const uint32_t peers = __match_any_sync(active, d);
            leader = __ffs(peers) - 1;
            below = __popc(peers & ((1u << lane) - 1u));
            if (lane == leader) { old = cnt[wid][d]; cnt[wid][d] = old + __popc(peers); }
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        lrank[r] = old + below;
        __syncwarp();
    }
    __syncthreads();
    {   // per digit: global tile offset + exclusive prefix over warps
        const int d = threadIdx.x;
        uint32_t run = offs[((int64_t)seg * tiles + tile) * 256 + d];
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) { const uint32_t c = cnt[w][d]; cnt[w][d] = run; run += c; }
    }
    __syncthreads();
    uint64_t* kout = keys_out + (int64_t)seg * seg_len;
#pragma unroll
    for (int r = 0; r < RS_KPT; r++) {
        const int64_t i = wbase + r * 32 + lane;
        if (i < seg_len) {
            const uint32_t d = (uint32_t)(key[r] >> shift) & 255u;
            const uint32_t pos = cnt[wid][d] + lrank[r];
            kout[pos] = key[r];
            if (HAS_VAL) vals_out[(int64_t)seg * seg_len + pos] = val[r];
        }
    }
}

__device__ __forceinline__ uint64_t order_key(double x) {
    const uint64_t b = (x == 0.0) ? 0ull : (uint64_t)__double_as_longlong(x);   // -0.0 ties with +0.0 under operator<
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// order keys: monotone map of IEEE doubles to unsigned integers; NaN raises a flag
__global__ void order_keys_kernel(const double* __restrict__ v, int64_t n, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                  int* __restrict__ nan_flag) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = v[i];
        if (x != x) *nan_flag = 1;
        keys[i] = order_key(x);
        vals[i] = (uint32_t)i;
    }
}

__global__ void widen_kernel(const uint32_t* __restrict__ vals, int64_t n, uint64_t* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = vals[i];
}

// ---- top-N selection (AbcSmc.cpp:645-646 keeps only the first N_pp entries of the order) ---------------------------------
// MSD radix select on the 64-bit order keys, 12 bits per level, two levels (sign + exponent, then 12 mantissa bits):
//   select_hist1_kernel : keys + level-1 histogram;
//   select_hist2_kernel : every CTA scans the level-1 histogram (16 KB) to find the bin b1 holding the top_n-th key,
//                         then histograms the next 12 bits of the keys inside b1;
//   select_compact_kernel: every CTA scans the level-2 histogram, then compacts (key, index) of all keys whose 24-bit
//                         prefix is <= the boundary prefix (the certain ones and the boundary bin);
//   select_rank_kernel  : the <= SEL_CAP candidates are ranked by (key, index) by counting (all pairs, one thread per candidate)
//                         and the first top_n indices written to their place. Ties come out in ascending particle index.
// If the boundary bin is too crowded (many identical distances) the caller falls back to the full radix sort.
constexpr int SEL_BINS = 4096;
constexpr int SEL_THREADS = 256;
constexpr int SEL_CAP = 16384;           // candidates the final CTA can sort (key 8 B + index 4 B each: 192 KB)

// block-wide: find the first bin whose inclusive prefix reaches `need` (1-based); returns bin and the count below it
__device__ __forceinline__ void find_bin(const uint32_t* __restrict__ hist, uint32_t need, uint32_t* sh, uint32_t& bin, uint32_t& below) {
    constexpr int PER = SEL_BINS / SEL_THREADS;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t c[PER], tot = 0;
#pragma unroll
    for (int j = 0; j < PER; j++) { c[j] = hist[tid * PER + j]; tot += c[j]; }
    uint32_t incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) sh[wid] = incl;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < wid; w++) base += sh[w];
    uint32_t run = base + incl - tot;
    __syncthreads();
    if (run < need && need <= run + tot) {          // exactly one thread
#pragma unroll
        for (int j = 0; j < PER; j++) {
            if (run < need && need <= run + c[j]) { sh[8] = (uint32_t)(tid * PER + j); sh[9] = run; }
            run += c[j];
        }
    }
    __syncthreads();
    bin = sh[8]; below = sh[9];
}

__global__ void __launch_bounds__(SEL_THREADS) select_hist1_kernel(const double* __restrict__ v, int64_t n, uint64_t* __restrict__ keys,
                                                                   uint32_t* __restrict__ hist1, int* __restrict__ nan_flag) {
    __shared__ uint32_t h[SEL_BINS];
    for (int i = threadIdx.x; i < SEL_BINS; i += SEL_THREADS) h[i] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * SEL_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * SEL_THREADS) {
        const double x = v[i];
        if (x != x) *nan_flag = 1;
        const uint64_t k = order_key(x);
        keys[i] = k;
        atomicAdd(&h[(uint32_t)(k >> 52)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SEL_BINS; i += SEL_THREADS) if (h[i]) atomicAdd(&hist1[i], h[i]);
}

// sel[0] = b1, sel[1] = #keys below bin b1
__global__ void __launch_bounds__(SEL_THREADS) select_hist2_kernel(const uint64_t* __restrict__ keys, int64_t n, uint32_t top_n,
                                                                   const uint32_t* __restrict__ hist1, uint32_t* __restrict__ hist2,
                                                                   uint32_t* __restrict__ sel) {
    __shared__ uint32_t h[SEL_BINS];
    __shared__ uint32_t sh[16];
    uint32_t b1, below;
    find_bin(hist1, top_n, sh, b1, below);
    if (blockIdx.x == 0 && threadIdx.x == 0) { sel[0] = b1; sel[1] = below; }
    for (int i = threadIdx.x; i < SEL_BINS; i += SEL_THREADS) h[i] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * SEL_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * SEL_THREADS) {
        const uint64_t k = keys[i];
        if ((uint32_t)(k >> 52) == b1) atomicAdd(&h[(uint32_t)(k >> 40) & (SEL_BINS - 1)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SEL_BINS; i += SEL_THREADS) if (h[i]) atomicAdd(&hist2[i], h[i]);
}

// cand_count[0] is the running number of candidates; candidates beyond `cap` are counted but not stored
__global__ void __launch_bounds__(SEL_THREADS) select_compact_kernel(const uint64_t* __restrict__ keys, int64_t n, uint32_t top_n,
                                                                     const uint32_t* __restrict__ hist2, const uint32_t* __restrict__ sel,
                                                                     uint32_t cap, uint64_t* __restrict__ cand_key,
                                                                     uint32_t* __restrict__ cand_idx, uint32_t* __restrict__ cand_count) {
    __shared__ uint32_t sh[16];
    const uint32_t b1 = sel[0], below1 = sel[1];
    uint32_t b2, below2;
    find_bin(hist2, top_n - below1, sh, b2, below2);
    const uint64_t limit = ((uint64_t)b1 << 12) | (uint64_t)b2;     // 24-bit boundary prefix
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * SEL_THREADS;
    const int64_t nround = (n + stride - 1) / stride;
    for (int64_t r = 0; r < nround; r++) {
        const int64_t i = r * stride + (int64_t)blockIdx.x * SEL_THREADS + threadIdx.x;
        const bool in = i < n;
        const uint64_t k = in ? keys[i] : ~0ull;
        const bool take = in && (k >> 40) <= limit;
        const uint32_t m = __ballot_sync(0xffffffffu, take);
        if (m) {
            uint32_t base = 0;
            if (lane == __ffs(m) - 1) base = atomicAdd(cand_count, (uint32_t)__popc(m));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
            if (take) {
                const uint32_t pos = base + __popc(m & ((1u << lane) - 1u));
                if (pos < cap) { cand_key[pos] = k; cand_idx[pos] = (uint32_t)i; }
            }
        }
    }
}

// Final order of the <= SEL_CAP candidates by counting: rank_i = #{j : (key_j, index_j) < (key_i, index_i)}, one thread per candidate,
// the candidates streamed through shared memory in tiles; the first top_n ranks are written to their place. All pairs on the
// whole GPU instead of one CTA's bitonic network (125 -> ~15 us for 6k candidates, 257 -> ~45 us for 12k). Indices are distinct,
// so ranks are a permutation: ties in the key come out in ascending particle index, as in the full sort.
constexpr int RK_T = 128;
__global__ void __launch_bounds__(RK_T) select_rank_kernel(const uint64_t* __restrict__ cand_key, const uint32_t* __restrict__ cand_idx,
                                                           const uint32_t* __restrict__ cand_count, uint32_t cap, uint32_t top_n,
                                                           uint64_t* __restrict__ order_out, int* __restrict__ overflow,
                                                           const int* __restrict__ nan_flag) {
    extern __shared__ __align__(16) unsigned char rk_smem[];
    const uint32_t total = *cand_count;
    const bool bad = total > cap || total < top_n;
    if (blockIdx.x == 0 && threadIdx.x == 0) { overflow[1] = *nan_flag; overflow[0] = bad ? 1 : 0; }   // the host reads (overflow, NaN flag) with one copy
    if (bad || blockIdx.x * RK_T >= total) return;
    // every candidate staged once (a tile at a time paid a global-load latency + two barriers per 128 comparisons: 0.3 ms)
    const uint32_t tot4 = (total + 3u) & ~3u;
    uint64_t* sk = (uint64_t*)rk_smem;
    uint32_t* si = (uint32_t*)(sk + tot4);
    for (uint32_t j = threadIdx.x; j < tot4; j += RK_T) {
        sk[j] = (j < total) ? cand_key[j] : ~0ull;       // padding sorts after everything
        si[j] = (j < total) ? cand_idx[j] : 0xffffffffu;
    }
    __syncthreads();
    const uint32_t i = blockIdx.x * RK_T + threadIdx.x;
    if (i >= total) return;
    const uint64_t ki = sk[i];
    const uint32_t ii = si[i];
    uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0;
    for (uint32_t j = 0; j < tot4; j += 4) {             // broadcast reads: every thread of the warp looks at the same candidates
        const ulonglong2 ka = *(const ulonglong2*)(sk + j), kb = *(const ulonglong2*)(sk + j + 2);
        const uint4 ia = *(const uint4*)(si + j);
        r0 += (ka.x < ki || (ka.x == ki && ia.x < ii)) ? 1u : 0u;
        r1 += (ka.y < ki || (ka.y == ki && ia.y < ii)) ? 1u : 0u;
        r2 += (kb.x < ki || (kb.x == ki && ia.z < ii)) ? 1u : 0u;
        r3 += (kb.y < ki || (kb.y == ki && ia.w < ii)) ? 1u : 0u;
    }
    const uint32_t rank = (r0 + r1) + (r2 + r3);
    if (rank < top_n) order_out[rank] = (uint64_t)ii;
}

}  // namespace

size_t radix_hist_bytes(int64_t seg_len, int n_seg) {
    const int64_t tiles = (seg_len + RS_TILE - 1) / RS_TILE;
    return align_up((size_t)n_seg * 256 * tiles * sizeof(uint32_t), 256);
}

int radix_sort_segments(abcb200_ctx* ctx, uint64_t* keys, uint64_t* keys_alt, uint32_t* vals, uint32_t* vals_alt, int64_t seg_len,
                        int n_seg, uint32_t* hist, const int* seg_valid) {
    if (seg_len <= 0 || n_seg <= 0) return ABCB200_OK;
    const int tiles = (int)((seg_len + RS_TILE - 1) / RS_TILE);
    uint64_t *kin = keys, *kout = keys_alt;
    uint32_t *vin = vals, *vout = vals_alt;
    for (int pass = 0; pass < 8; pass++) {
        const int shift = pass * 8;
        LAUNCH(ctx, radix_hist_kernel, dim3(tiles, n_seg), RS_THREADS, 0, kin, seg_len, tiles, shift, hist, seg_valid);
        LAUNCH(ctx, radix_scan_kernel, n_seg, 256, 0, hist, tiles, seg_valid);
        if (vals) LAUNCH(ctx, radix_scatter_kernel<true>, dim3(tiles, n_seg), RS_THREADS, 0, kin, kout, vin, vout, seg_len, tiles, shift, hist, seg_valid);
        else LAUNCH(ctx, radix_scatter_kernel<false>, dim3(tiles, n_seg), RS_THREADS, 0, kin, kout, vin, vout, seg_len, tiles, shift, hist, seg_valid);
        uint64_t* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    return ABCB200_OK;   // 8 passes: the sorted data are back in keys / vals
}

size_t order_ws_bytes(int64_t n) {
    return 2 * align_up((size_t)n * 8, 256) + 2 * align_up((size_t)n * 4, 256) + radix_hist_bytes(n, 1) + align_up((size_t)SEL_CAP * 12, 256) +
           align_up((2 * SEL_BINS + 16) * 4, 256) + 2048;
}

// top_n << n: radix select + small sort. Returns 1 in *done_host when the result was produced, 0 when the caller must run the full sort.
static int select_dev(abcb200_ctx* ctx, const double* v, int64_t n, int64_t top_n, uint64_t* order_out, uint64_t* keys, int* flag,
                      bool* done_host) {
    uint32_t* meta = ws_new<uint32_t>(ctx, 2 * SEL_BINS + 16);     // hist1, hist2, sel[2], cand_count, overflow
    uint64_t* cand_key = ws_new<uint64_t>(ctx, SEL_CAP);
    uint32_t* cand_idx = ws_new<uint32_t>(ctx, SEL_CAP);
    if (!meta || !cand_key || !cand_idx) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in order_dev (select)");
    uint32_t *hist1 = meta, *hist2 = meta + SEL_BINS, *sel = meta + 2 * SEL_BINS, *cand_count = sel + 4;
    int* overflow = (int*)(sel + 5);
    CUDA_TRY(ctx, cudaMemsetAsync(meta, 0, (2 * SEL_BINS + 16) * sizeof(uint32_t), ctx->stream));
    const int grid = (int)max((int64_t)1, min((n + SEL_THREADS - 1) / SEL_THREADS, (int64_t)(2 * ctx->sm_count)));
    LAUNCH(ctx, select_hist1_kernel, grid, SEL_THREADS, 0, v, n, keys, hist1, flag);
    LAUNCH(ctx, select_hist2_kernel, grid, SEL_THREADS, 0, keys, n, (uint32_t)top_n, hist1, hist2, sel);
    LAUNCH(ctx, select_compact_kernel, grid, SEL_THREADS, 0, keys, n, (uint32_t)top_n, hist2, sel, (uint32_t)SEL_CAP, cand_key, cand_idx, cand_count);
    const size_t rk_smem = (size_t)SEL_CAP * 12;
    CUDA_TRY(ctx, cudaFuncSetAttribute(select_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rk_smem));
    LAUNCH(ctx, select_rank_kernel, SEL_CAP / RK_T, RK_T, rk_smem, cand_key, cand_idx, cand_count, (uint32_t)SEL_CAP, (uint32_t)top_n, order_out, overflow, (const int*)flag);
    ABC_TRY(hpin_reserve(ctx, 64));
    int* h = (int*)ctx->hpin;
    CUDA_TRY(ctx, cudaMemcpyAsync(h, overflow, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));     // [0] overflow, [1] NaN flag
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (h[1]) ABC_FAIL(ctx, ABCB200_ENAN, "NaN among the %lld values to order (std::sort comparator would be inconsistent)", (long long)n);
    *done_host = h[0] == 0;
    return ABCB200_OK;
}

int order_dev(abcb200_ctx* ctx, const double* v, int64_t n, int64_t top_n, uint64_t* order_out) {
    if (n <= 0) return ABCB200_OK;
    if (n > 0xffffffffll) ABC_FAIL(ctx, ABCB200_EINVAL, "ordered: n=%lld exceeds 2^32", (long long)n);
    if (top_n <= 0 || top_n > n) top_n = n;
    uint64_t* keys = ws_new<uint64_t>(ctx, n);
    uint64_t* keys_alt = ws_new<uint64_t>(ctx, n);
    uint32_t* vals = ws_new<uint32_t>(ctx, n);
    uint32_t* vals_alt = ws_new<uint32_t>(ctx, n);
    uint32_t* hist = (uint32_t*)ws_alloc(ctx, radix_hist_bytes(n, 1));
    int* flag = ws_new<int>(ctx, 1);
    if (!keys || !keys_alt || !vals || !vals_alt || !hist || !flag) ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace exhausted in order_dev");
    CUDA_TRY(ctx, cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
    if (top_n + 4096 <= SEL_CAP && n >= 4 * top_n && n >= 4096) {
        bool done = false;
        ABC_TRY(select_dev(ctx, v, n, top_n, order_out, keys, flag, &done));
        if (done) return ABCB200_OK;     // else: boundary bin too crowded (massive ties) -> full sort below
    }
    const int grid = (int)max((int64_t)1, min((n + 255) / 256, (int64_t)(8 * ctx->sm_count)));
    LAUNCH(ctx, order_keys_kernel, grid, 256, 0, v, n, keys, vals, flag);
    ABC_TRY(radix_sort_segments(ctx, keys, keys_alt, vals, vals_alt, n, 1, hist, nullptr));
    LAUNCH(ctx, widen_kernel, (int)max((int64_t)1, min((top_n + 255) / 256, (int64_t)(8 * ctx->sm_count))), 256, 0, vals, top_n, order_out);
    ABC_TRY(hpin_reserve(ctx, 64));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hpin, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (*(int*)ctx->hpin) ABC_FAIL(ctx, ABCB200_ENAN, "NaN among the %lld values to order (std::sort comparator would be inconsistent)", (long long)n);
    return ABCB200_OK;
}
