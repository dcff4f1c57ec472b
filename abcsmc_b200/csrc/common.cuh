// common.cuh — context, workspace arena, launch accounting and device helpers shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/abcsmc_b200.h"

#define ABC_NSTAGES ABCB200_NSTAGES
#define ABC_NKERNELS ABCB200_NKERNELS
#define ABC_K_LOO 9   // kernel timer of the batched leave-one-out refits

struct SmPartition {     // one green-context split of the device's SMs
    int state;           // 0 not tried, 1 live, -1 unavailable
    int want, small_sms, rest_sms;
    cudaStream_t small, rest;
    void* green[2];      // CUgreenCtx handles (kept alive for the context's lifetime)
};

struct abcb200_ctx {
    int device;
    int sm_count;
    cudaStream_t stream;
    cudaStream_t own_stream;
    char* ws;            // device workspace arena (bump allocated, reset per API call)
    size_t ws_cap, ws_off;
    char* hpin;          // pinned host scratch for small results
    size_t hpin_cap;
    uint64_t launches;
    uint64_t exact_tests; // signed-rank tests that needed exact ranks (diagnostic)
    uint64_t exact_radix_calls;         // selections whose exact tests went through the radix sort (forced, or the fine-bin level gave up)
    uint64_t stat_tests, stat_level2;   // last selection: tests in total (sum of ref_y) and tests that reached level 2
    int stage_timers;                   // per-stage CUDA events on / off
    uint32_t kernel_timers;             // bit k: CUDA-event bracket of hot kernel k
    uint64_t stat_pipe_block;           // components per block of the last pipelined fit + hold-out (0: stage after stage)
    uint64_t stat_pls_loop;             // component loop of the last fit: 1 pls_defl_kernel (all on chip), 2 pls_gram_kernel, 3 pls_wide.cu
    char err[512];
    cudaEvent_t ev[ABC_NSTAGES][2];
    bool ev_valid[ABC_NSTAGES];
    cudaEvent_t kev[ABC_NKERNELS][2];   // CUDA-event brackets of individual hot kernels (roofline reporting)
    bool kev_valid[ABC_NKERNELS];
    int smem_optin;      // max dynamic shared memory per block
    // Lanes of the pipelined fit (api.cu: rank_fit_holdout_pipelined; context.cu: ctx_lanes): a producer stream and a consumer stream,
    // either ordinary streams with high / low priority or the two sides of an SM partition (green contexts, created on first use).
    cudaStream_t prio_small, prio_rest;
    SmPartition part[2];
    int last_partition;  // SMs of the producer side of the lanes handed out last (0: ordinary streams)
    cudaEvent_t pev[24]; // cross-stream dependencies of the pipelined fit
    int tie_order;              // placement of exact distance ties: 0 ascending particle index, 1 as libstdc++'s std::sort leaves them (api.cu: tie_order_*)
    uint64_t stat_tie_resorts;  // rankings whose order was re-derived on the host because exact ties reached the output (mode 1)
    cudaStream_t copy_stream;   // H2D of the host entry points, in column blocks, so that S1 starts on the blocks that have arrived
    cudaEvent_t cev[12];
};

bool ctx_lanes(abcb200_ctx* ctx, int nsmall, cudaStream_t* small, cudaStream_t* rest);

// Launches inside the scope go to stream `s` (LAUNCH and the timers read ctx->stream).
struct StreamScope {
    abcb200_ctx* c; cudaStream_t saved;
    StreamScope(abcb200_ctx* ctx, cudaStream_t s) : c(ctx), saved(ctx->stream) { ctx->stream = s; }
    ~StreamScope() { c->stream = saved; }
};

#define CUDA_TRY(ctx, call)                                                                              \
    do {                                                                                                 \
        cudaError_t _e = (call);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d: %s -> %s", __FILE__, __LINE__, #call,       \
                     cudaGetErrorString(_e));                                                            \
            return ABCB200_ECUDA;                                                                        \
        }                                                                                                \
    } while (0)

#define ABC_TRY(call)                 \
    do {                              \
        int _r = (call);              \
        if (_r != ABCB200_OK) return _r; \
    } while (0)

#define ABC_FAIL(ctx, code, ...)                                  \
    do {                                                          \
        snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__);    \
        return (code);                                            \
    } while (0)

// Every kernel launch goes through this macro so that launches are counted and launch errors surface.
#define LAUNCH(ctx, kernel, grid, block, smem, ...)                                  \
    do {                                                                             \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);             \
        (ctx)->launches++;                                                           \
        CUDA_TRY(ctx, cudaGetLastError());                                           \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Workspace arena. ws_reserve may reallocate (synchronises); ws_alloc never fails after a sufficient reserve.
int ws_reserve(abcb200_ctx* ctx, size_t bytes);
void* ws_alloc(abcb200_ctx* ctx, size_t bytes);
static inline void ws_reset(abcb200_ctx* ctx) { ctx->ws_off = 0; }
template <typename T> static inline T* ws_new(abcb200_ctx* ctx, size_t n) { return (T*)ws_alloc(ctx, n * sizeof(T)); }
int hpin_reserve(abcb200_ctx* ctx, size_t bytes);

// CUDA-event instrumentation (abcb200_set_timers): ctx->stage_timers switches the per-stage brackets, bit k of
// ctx->kernel_timers the bracket of hot kernel k. Every record costs launch path (~4 us each at the C2 shape: 20 records were
// 14 % of the step), so everything is off by default; bench.py switches on what it reports.
static inline void stage_begin(abcb200_ctx* ctx, int s) { if (ctx->stage_timers) cudaEventRecord(ctx->ev[s][0], ctx->stream); }
static inline void stage_end(abcb200_ctx* ctx, int s) { if (ctx->stage_timers) { cudaEventRecord(ctx->ev[s][1], ctx->stream); ctx->ev_valid[s] = true; } }
static inline void kernel_begin(abcb200_ctx* ctx, int k) { if ((ctx->kernel_timers >> k) & 1u) cudaEventRecord(ctx->kev[k][0], ctx->stream); }
static inline void kernel_end(abcb200_ctx* ctx, int k) { if ((ctx->kernel_timers >> k) & 1u) { cudaEventRecord(ctx->kev[k][1], ctx->stream); ctx->kev_valid[k] = true; } }

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// 32 warp sums at once: on return v[0] of lane l holds the sum over the warp of the callers' v[l]. Same pairing tree as
// warp_sum (xor 16, 8, 4, 2, 1), so the bits are identical, with 31 double shuffles instead of 160.
__device__ __forceinline__ double warp_sum32_transposed(double (&v)[32]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int j = 0; j < off; j++) {
            const double send = upper ? v[j] : v[j + off];
            const double keep = upper ? v[j + off] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Deterministic block-wide sum (fixed shuffle tree + fixed cross-warp order). `red` holds >= 32 doubles.
// The result is returned to every thread.
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double r = (lane < nw) ? red[lane] : 0.0;
    r = warp_sum(r);
    return r;
}

// FP64 tensor-core tile product: D(8x8) += A(8x4, row) * B(4x8, col). SASS: DMMA.8x8x4.
// Fragment ownership (PTX ISA, mma.m8n8k4 .f64): a = A[lane>>2][lane&3]; b = B[lane&3][lane>>2];
// c0,c1 = C[lane>>2][2*(lane&3) + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// exp(-q) for q >= 0 (clamped), ~1e-15 relative: n = rint(-q*log2 e); r = -q - n*ln2 (two-part);
// degree-12 polynomial; scale by 2^n through the exponent field; results below 2^-1021 flush to 0.
__device__ __forceinline__ double exp_neg(double q) {
    const double x = -q;
    const double L2E = 1.4426950408889634074, LN2HI = 6.93147180369123816490e-01, LN2LO = 1.90821492927058770002e-10;
    const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52
    double t = fma(x, L2E, MAGIC);
    int n = __double2loint(t);
    double fn = t - MAGIC;
    double r = fma(fn, -LN2HI, x);
    r = fma(fn, -LN2LO, r);
    double p = 2.08767569878681e-09;            // 1/12!
    p = fma(p, r, 2.505210838544172e-08);       // 1/11!
    p = fma(p, r, 2.755731922398589e-07);
    p = fma(p, r, 2.755731922398589e-06);
    p = fma(p, r, 2.480158730158730e-05);
    p = fma(p, r, 1.984126984126984e-04);
    p = fma(p, r, 1.388888888888889e-03);
    p = fma(p, r, 8.333333333333333e-03);
    p = fma(p, r, 4.166666666666666e-02);
    p = fma(p, r, 1.666666666666667e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    int hi = __double2hiint(p) + (n << 20);
    double res = __hiloint2double(hi, __double2loint(p));
    return (q < 708.0) ? res : 0.0;   // also maps NaN -> 0? no: NaN compares false -> 0; callers pre-check NaN
}

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers ------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

#endif  // __CUDACC__
