// context.cu — context lifetime, workspace arena, pinned scratch, stage timers.
#include <stdlib.h>

#include "common.cuh"

#include <cuda.h>

#include <new>

namespace {
// Green contexts (CUDA >= 12.4 driver API, resolved through the runtime: no link-time dependency on libcuda): an 8-SM
// partition for the latency-bound one-CTA component loop and the remaining SMs for the kernels that run beside it.
// Measured with tools/green_probe.cu: ten back-to-back one-CTA launches next to saturating kernels take 1.01 ms on their own
// partition against 1.87 ms on a high-priority ordinary stream (each launch waits for an SM to drain) and 0.99 ms alone.
typedef CUresult (*pfnGetRes)(CUdevice, CUdevResource*, CUdevResourceType);
typedef CUresult (*pfnSplit)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
typedef CUresult (*pfnDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int);
typedef CUresult (*pfnCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
typedef CUresult (*pfnStream)(CUstream*, CUgreenCtx, unsigned int, int);
typedef CUresult (*pfnDestroy)(CUgreenCtx);
template <class T> T drv_entry(const char* name) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); return nullptr; }
    return (T)p;
}
#define PART_FAIL(what, code)                                                                                   \
    do {                                                                                                        \
        if (getenv("ABCB200_DEBUG")) fprintf(stderr, "[abcb200] no SM partition: %s failed (%d)\n", what, (int)(code)); \
        return false;                                                                                           \
    } while (0)
bool make_partition(abcb200_ctx* ctx, int nsmall, SmPartition* out) {
    auto getRes = drv_entry<pfnGetRes>("cuDeviceGetDevResource"); auto split = drv_entry<pfnSplit>("cuDevSmResourceSplitByCount");
    auto mkDesc = drv_entry<pfnDesc>("cuDevResourceGenerateDesc"); auto mkCtx = drv_entry<pfnCreate>("cuGreenCtxCreate");
    auto mkStream = drv_entry<pfnStream>("cuGreenCtxStreamCreate"); auto rmCtx = drv_entry<pfnDestroy>("cuGreenCtxDestroy");
    if (!getRes || !split || !mkDesc || !mkCtx || !mkStream || !rmCtx) PART_FAIL("cudaGetDriverEntryPoint(green-context API)", 0);
    CUdevResource all, small, rest;
    unsigned int ng = 1;
    CUresult r;
    if ((r = getRes((CUdevice)ctx->device, &all, CU_DEV_RESOURCE_TYPE_SM)) != CUDA_SUCCESS) PART_FAIL("cuDeviceGetDevResource", r);
    if ((r = split(&small, &ng, &all, &rest, 0, (unsigned)nsmall)) != CUDA_SUCCESS || ng != 1 || rest.sm.smCount < 64) PART_FAIL("cuDevSmResourceSplitByCount", r);
    CUdevResourceDesc dA, dB;
    if ((r = mkDesc(&dA, &small, 1)) != CUDA_SUCCESS || (r = mkDesc(&dB, &rest, 1)) != CUDA_SUCCESS) PART_FAIL("cuDevResourceGenerateDesc", r);
    CUgreenCtx gA = nullptr, gB = nullptr;
    if ((r = mkCtx(&gA, dA, (CUdevice)ctx->device, CU_GREEN_CTX_DEFAULT_STREAM)) != CUDA_SUCCESS) PART_FAIL("cuGreenCtxCreate (small)", r);
    if ((r = mkCtx(&gB, dB, (CUdevice)ctx->device, CU_GREEN_CTX_DEFAULT_STREAM)) != CUDA_SUCCESS) { rmCtx(gA); PART_FAIL("cuGreenCtxCreate (rest)", r); }
    CUstream sA = nullptr, sB = nullptr;
    if ((r = mkStream(&sA, gA, CU_STREAM_NON_BLOCKING, 0)) != CUDA_SUCCESS || (r = mkStream(&sB, gB, CU_STREAM_NON_BLOCKING, 0)) != CUDA_SUCCESS) {
        if (sA) cudaStreamDestroy((cudaStream_t)sA);
        rmCtx(gA); rmCtx(gB);
        PART_FAIL("cuGreenCtxStreamCreate", r);
    }
    out->small = (cudaStream_t)sA; out->rest = (cudaStream_t)sB;
    out->green[0] = gA; out->green[1] = gB;
    out->small_sms = (int)small.sm.smCount; out->rest_sms = (int)rest.sm.smCount;
    if (getenv("ABCB200_DEBUG")) fprintf(stderr, "[abcb200] SM partition %d + %d\n", out->small_sms, out->rest_sms);
    return true;
}
}  // namespace

extern "C" int abcb200_create(int device, abcb200_ctx** out) {
    if (!out) return ABCB200_EINVAL;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return ABCB200_ENODEV;
    if (cudaSetDevice(device) != cudaSuccess) return ABCB200_ENODEV;
    abcb200_ctx* ctx = new (std::nothrow) abcb200_ctx();
    if (!ctx) return ABCB200_ENOMEM;
    memset(ctx, 0, sizeof(*ctx));
    if (const char* t = getenv("ABCB200_TIMERS")) { ctx->stage_timers = atoi(t) & 1; ctx->kernel_timers = (atoi(t) & 2) ? 0xffffffffu : 0u; }   // debug: 1 stages, 2 kernels, 3 both
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return ABCB200_ENODEV; }
    if (prop.major < 10) {   // sm_100a cubin only: fail loudly, there is no other code path
        delete ctx;
        return ABCB200_ENODEV;
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return ABCB200_ECUDA; }
    ctx->stream = ctx->own_stream;
    {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaStreamCreateWithPriority(&ctx->prio_small, cudaStreamNonBlocking, hi) != cudaSuccess ||
            cudaStreamCreateWithPriority(&ctx->prio_rest, cudaStreamNonBlocking, lo) != cudaSuccess) { delete ctx; return ABCB200_ECUDA; }
    }
    for (auto& e : ctx->pev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return ABCB200_ECUDA; }
    for (auto& e : ctx->cev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    for (int s = 0; s < ABC_NSTAGES; s++) {
        cudaEventCreate(&ctx->ev[s][0]);
        cudaEventCreate(&ctx->ev[s][1]);
    }
    for (int k = 0; k < ABC_NKERNELS; k++) {
        cudaEventCreate(&ctx->kev[k][0]);
        cudaEventCreate(&ctx->kev[k][1]);
    }
    *out = ctx;
    return ABCB200_OK;
}

extern "C" int abcb200_destroy(abcb200_ctx* ctx) {
    if (!ctx) return ABCB200_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->hpin) cudaFreeHost(ctx->hpin);
    for (int s = 0; s < ABC_NSTAGES; s++) { cudaEventDestroy(ctx->ev[s][0]); cudaEventDestroy(ctx->ev[s][1]); }
    for (int k = 0; k < ABC_NKERNELS; k++) { cudaEventDestroy(ctx->kev[k][0]); cudaEventDestroy(ctx->kev[k][1]); }
    cudaStreamSynchronize(ctx->prio_small); cudaStreamSynchronize(ctx->prio_rest);
    for (auto& e : ctx->pev) cudaEventDestroy(e);
    cudaStreamSynchronize(ctx->copy_stream);
    for (auto& e : ctx->cev) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->prio_small); cudaStreamDestroy(ctx->prio_rest);
    for (auto& pt : ctx->part) {
        if (pt.state != 1) continue;
        cudaStreamSynchronize(pt.small); cudaStreamSynchronize(pt.rest);
        cudaStreamDestroy(pt.small); cudaStreamDestroy(pt.rest);
        if (auto rmCtx = drv_entry<pfnDestroy>("cuGreenCtxDestroy")) { rmCtx((CUgreenCtx)pt.green[0]); rmCtx((CUgreenCtx)pt.green[1]); }
    }
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return ABCB200_OK;
}

// Two streams for a producer / consumer pair of kernel sequences. nsmall == 0: ordinary streams, high priority for the producer.
// nsmall > 0: an SM partition (green contexts, created on first use and kept): `nsmall` SMs (a multiple of 8) for the producer, the
// rest for the consumers. Returns false when a partition was asked for and cannot be had (ABCB200_SM_PARTITION=0 refuses them all).
bool ctx_lanes(abcb200_ctx* ctx, int nsmall, cudaStream_t* small, cudaStream_t* rest) {
    if (nsmall <= 0) { *small = ctx->prio_small; *rest = ctx->prio_rest; ctx->last_partition = 0; return true; }
    if (const char* e = getenv("ABCB200_SM_PARTITION")) if (!atoi(e)) return false;
    for (auto& pt : ctx->part) {
        if (pt.state == 0) { pt.want = nsmall; pt.state = make_partition(ctx, nsmall, &pt) ? 1 : -1; }
        if (pt.want != nsmall) continue;
        if (pt.state != 1) return false;
        *small = pt.small; *rest = pt.rest; ctx->last_partition = pt.small_sms;
        return true;
    }
    return false;
}

extern "C" int abcb200_set_stream(abcb200_ctx* ctx, void* cuda_stream) {
    if (!ctx) return ABCB200_EINVAL;
    cudaStreamSynchronize(ctx->stream);
    // NULL is CUDA's legacy default stream (a valid choice, e.g. torch's default); ABCB200_OWN_STREAM restores the context's own
    ctx->stream = (cuda_stream == ABCB200_OWN_STREAM) ? ctx->own_stream : (cudaStream_t)cuda_stream;
    return ABCB200_OK;
}

extern "C" int abcb200_synchronize(abcb200_ctx* ctx) {
    if (!ctx) return ABCB200_EINVAL;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ABCB200_OK;
}

extern "C" int abcb200_set_timers(abcb200_ctx* ctx, int stage_on, uint32_t kernel_mask) {
    if (!ctx) return ABCB200_EINVAL;
    ctx->stage_timers = stage_on ? 1 : 0;
    ctx->kernel_timers = kernel_mask;
    if (!stage_on) for (auto& v : ctx->ev_valid) v = false;
    for (int k = 0; k < ABC_NKERNELS; k++) if (!((kernel_mask >> k) & 1u)) ctx->kev_valid[k] = false;
    return ABCB200_OK;
}

extern "C" const char* abcb200_last_error(abcb200_ctx* ctx) { return ctx ? ctx->err : "null context"; }
extern "C" uint64_t abcb200_launch_count(abcb200_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" uint64_t abcb200_exact_test_count(abcb200_ctx* ctx) { return ctx ? ctx->exact_tests : 0; }
extern "C" int abcb200_set_tie_order(abcb200_ctx* ctx, int mode) {
    if (!ctx || (mode != 0 && mode != 1)) return ABCB200_EINVAL;
    ctx->tie_order = mode;
    return ABCB200_OK;
}

extern "C" uint64_t abcb200_stat(abcb200_ctx* ctx, int which) {
    if (!ctx) return 0;
    switch (which) {
        case 0: return ctx->launches;
        case 1: return ctx->stat_tests;
        case 2: return ctx->stat_level2;
        case 3: return ctx->exact_tests;
        case 4: return ctx->stat_pls_loop;
        case 5: return ctx->stat_pipe_block;
        case 6: return (uint64_t)ctx->last_partition;
        case 7: return ctx->exact_radix_calls;
        case 8: return ctx->stat_tie_resorts;
        default: return 0;
    }
}

extern "C" int abcb200_host_alloc(size_t bytes, void** out) {
    if (!out) return ABCB200_EINVAL;
    return cudaHostAlloc(out, bytes, cudaHostAllocDefault) == cudaSuccess ? ABCB200_OK : ABCB200_ENOMEM;
}
extern "C" int abcb200_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? ABCB200_OK : ABCB200_ECUDA; }

extern "C" double abcb200_stage_ms(abcb200_ctx* ctx, int stage) {
    if (!ctx || stage < 0 || stage >= ABC_NSTAGES) return -1.0;
    if (!ctx->ev_valid[stage]) return 0.0;
    float ms = 0;
    if (cudaEventSynchronize(ctx->ev[stage][1]) != cudaSuccess) return -1.0;
    if (cudaEventElapsedTime(&ms, ctx->ev[stage][0], ctx->ev[stage][1]) != cudaSuccess) return -1.0;
    return (double)ms;
}

extern "C" double abcb200_kernel_ms(abcb200_ctx* ctx, int kernel) {
    if (!ctx || kernel < 0 || kernel >= ABC_NKERNELS) return -1.0;
    if (!ctx->kev_valid[kernel]) return 0.0;
    float ms = 0;
    if (cudaEventSynchronize(ctx->kev[kernel][1]) != cudaSuccess) return -1.0;
    if (cudaEventElapsedTime(&ms, ctx->kev[kernel][0], ctx->kev[kernel][1]) != cudaSuccess) return -1.0;
    return (double)ms;
}

int ws_reserve(abcb200_ctx* ctx, size_t bytes) {
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    ctx->ws_off = 0;
    if (bytes <= ctx->ws_cap) return ABCB200_OK;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->ws) { cudaFree(ctx->ws); ctx->ws = nullptr; ctx->ws_cap = 0; }
    size_t cap = align_up(bytes + (bytes >> 3), (size_t)1 << 21);
    if (cudaMalloc(&ctx->ws, cap) != cudaSuccess) {
        cudaGetLastError();
        ABC_FAIL(ctx, ABCB200_ENOMEM, "workspace of %zu bytes could not be allocated", cap);
    }
    ctx->ws_cap = cap;
    return ABCB200_OK;
}

void* ws_alloc(abcb200_ctx* ctx, size_t bytes) {
    size_t off = align_up(ctx->ws_off, 256);
    if (off + bytes > ctx->ws_cap) return nullptr;   // callers reserve an upper bound first
    ctx->ws_off = off + bytes;
    return ctx->ws + off;
}

int hpin_reserve(abcb200_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->hpin_cap) return ABCB200_OK;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->hpin) cudaFreeHost(ctx->hpin);
    ctx->hpin = nullptr; ctx->hpin_cap = 0;
    size_t cap = align_up(bytes, 4096);
    if (cudaHostAlloc((void**)&ctx->hpin, cap, cudaHostAllocDefault) != cudaSuccess) ABC_FAIL(ctx, ABCB200_ENOMEM, "pinned scratch of %zu bytes failed", cap);
    ctx->hpin_cap = cap;
    return ABCB200_OK;
}
