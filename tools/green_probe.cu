// green_probe.cu — does SM partitioning with green contexts (cuGreenCtxCreate) let a one-CTA latency-bound kernel run BESIDE
// whole-GPU throughput kernels launched with the runtime API? Measures: (a) spin kernel alone, (b) spin kernel on partition A while
// saturating kernels run on partition B, (c) the same on two ordinary streams (no partition), (d) chunked spin launches (10 x) against
// a stream of saturating kernels, with and without partitions. Build: nvcc -arch=sm_100a -o green_probe green_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void __launch_bounds__(512, 1) spin_kernel(long long cycles, long long* out) {
    extern __shared__ double big[];
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) { big[threadIdx.x] += 1.0; }
    if (threadIdx.x == 0) out[0] = clock64() - t0;
}
__global__ void __launch_bounds__(128, 3) busy_kernel(const double* __restrict__ x, double* __restrict__ y, long long n, int reps) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double v = x[i];
        for (int r = 0; r < reps; r++) v = fma(v, 1.0000001, 1e-9);
        y[i] = v;
    }
}

typedef CUresult (*pfnGetRes)(CUdevice, CUdevResource*, CUdevResourceType);
typedef CUresult (*pfnSplit)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
typedef CUresult (*pfnDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int);
typedef CUresult (*pfnCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
typedef CUresult (*pfnStream)(CUstream*, CUgreenCtx, unsigned int, int);

template <class T> static T entry(const char* name) {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { printf("no driver entry %s\n", name); return nullptr; }
    return (T)p;
}

int main() {
    CK(cudaSetDevice(0));
    CK(cudaFree(0));
    cudaStream_t sA = nullptr, sB = nullptr, nA, nB;
    int lo, hi; CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&nA, cudaStreamNonBlocking, hi)); CK(cudaStreamCreateWithPriority(&nB, cudaStreamNonBlocking, lo));
    auto getRes = entry<pfnGetRes>("cuDeviceGetDevResource"); auto split = entry<pfnSplit>("cuDevSmResourceSplitByCount");
    auto mkDesc = entry<pfnDesc>("cuDevResourceGenerateDesc"); auto mkCtx = entry<pfnCreate>("cuGreenCtxCreate"); auto mkStream = entry<pfnStream>("cuGreenCtxStreamCreate");
    if (getRes && split && mkDesc && mkCtx && mkStream) {
        CUdevResource all, small, rest; unsigned int ng = 1; CUresult r;
        r = getRes(0, &all, CU_DEV_RESOURCE_TYPE_SM); printf("getRes %d sm=%u\n", (int)r, all.sm.smCount);
        r = split(&small, &ng, &all, &rest, 0, 8); printf("split %d groups=%u small=%u rest=%u\n", (int)r, ng, small.sm.smCount, rest.sm.smCount);
        CUdevResourceDesc dA, dB; CUgreenCtx gA, gB;
        r = mkDesc(&dA, &small, 1); printf("descA %d\n", (int)r); r = mkDesc(&dB, &rest, 1); printf("descB %d\n", (int)r);
        r = mkCtx(&gA, dA, 0, CU_GREEN_CTX_DEFAULT_STREAM); printf("ctxA %d\n", (int)r); r = mkCtx(&gB, dB, 0, CU_GREEN_CTX_DEFAULT_STREAM); printf("ctxB %d\n", (int)r);
        r = mkStream((CUstream*)&sA, gA, CU_STREAM_NON_BLOCKING, 0); printf("streamA %d\n", (int)r); r = mkStream((CUstream*)&sB, gB, CU_STREAM_NON_BLOCKING, 0); printf("streamB %d\n", (int)r);
    }
    const long long n = 1ll << 26;
    double *x, *y; long long* out;
    CK(cudaMalloc(&x, n * 8)); CK(cudaMalloc(&y, n * 8)); CK(cudaMalloc(&out, 64)); CK(cudaMemset(x, 0, n * 8));
    CK(cudaFuncSetAttribute(busy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024));   // 3 x 40 KB per SM: the 200 KB spin CTA cannot co-reside
    CK(cudaFuncSetAttribute(spin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaEvent_t e0, e1, eb0, eb1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&eb0)); CK(cudaEventCreate(&eb1));
    const long long cyc = 1900000;   // ~1 ms
    auto run = [&](const char* name, cudaStream_t a, cudaStream_t b, int chunks, int nbusy) {
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0, a)); if (b) CK(cudaEventRecord(eb0, b));
            for (int c = 0; c < chunks; c++) {
                spin_kernel<<<1, 512, 200 * 1024, a>>>(cyc / chunks, out);
                if (b) for (int k = 0; k < nbusy / chunks; k++) busy_kernel<<<148 * 12, 128, 40 * 1024, b>>>(x, y, n, 64);
            }
            CK(cudaEventRecord(e1, a)); if (b) CK(cudaEventRecord(eb1, b));
            CK(cudaDeviceSynchronize());
            float ma = 0, mb = 0, tot = 0; CK(cudaEventElapsedTime(&ma, e0, e1));
            if (b) { CK(cudaEventElapsedTime(&mb, eb0, eb1)); CK(cudaEventElapsedTime(&tot, e0, eb1)); }
            if (rep == 1) printf("%-44s spin stream %.3f ms, busy stream %.3f ms, start A -> end B %.3f ms\n", name, ma, mb, tot);
        }
        CK(cudaGetLastError());
    };
    run("spin alone (normal stream)", nA, nullptr, 1, 0);
    run("spin x10 chunks alone", nA, nullptr, 10, 0);
    {   // busy alone
        CK(cudaDeviceSynchronize()); CK(cudaEventRecord(eb0, nB));
        for (int k = 0; k < 10; k++) busy_kernel<<<148 * 12, 128, 40 * 1024, nB>>>(x, y, n, 64);
        CK(cudaEventRecord(eb1, nB)); CK(cudaDeviceSynchronize()); float mb; CK(cudaEventElapsedTime(&mb, eb0, eb1)); printf("busy x10 alone (148 SMs) %.3f ms\n", mb);
    }
    run("normal streams (prio), 1 spin + 10 busy", nA, nB, 1, 10);
    run("normal streams (prio), 10 spin chunks + 10 busy", nA, nB, 10, 10);
    if (sA && sB) {
        run("spin alone on green A", sA, nullptr, 1, 0);
        {
            CK(cudaDeviceSynchronize()); CK(cudaEventRecord(eb0, sB));
            for (int k = 0; k < 10; k++) busy_kernel<<<148 * 12, 128, 40 * 1024, sB>>>(x, y, n, 64);
            CK(cudaEventRecord(eb1, sB)); CK(cudaDeviceSynchronize()); float mb; CK(cudaEventElapsedTime(&mb, eb0, eb1)); printf("busy x10 alone on green B %.3f ms\n", mb);
        }
        run("green A/B, 1 spin + 10 busy", sA, sB, 1, 10);
        run("green A/B, 10 spin chunks + 10 busy", sA, sB, 10, 10);
        // cross-stream event dependency between green streams and a normal stream
        CK(cudaEventRecord(e0, nA)); CK(cudaStreamWaitEvent(sA, e0, 0)); spin_kernel<<<1, 512, 200 * 1024, sA>>>(1000, out); CK(cudaEventRecord(e1, sA)); CK(cudaStreamWaitEvent(nA, e1, 0));
        CK(cudaDeviceSynchronize()); printf("cross-stream events ok\n");
    }
    return 0;
}
