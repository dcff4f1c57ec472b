// lat_probe.cu — dependent-chain latencies that bound the single-CTA PLS component loop (cycles, one SM).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lat_probe tools/lat_probe.cu && ./tools/lat_probe
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void probe(long long* out, double* sink, int nthreads_sync) {
    __shared__ double sm[1024];
    const int tid = threadIdx.x;
    sm[tid % 1024] = tid * 1e-3;
    __syncthreads();
    double a = 1.0 + tid * 1e-9, b = 1.0 - tid * 1e-9, c0 = 0, c1 = 0;
    long long t0, t1;
    const int N = 512;
    // 1. dependent DMMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) dmma(c0, c1, a, b);
    t1 = clock64();
    if (tid == 0) out[0] = (t1 - t0) / N;
    // 2. dependent DFMA chain
    double f = a;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) f = fma(f, b, a);
    t1 = clock64();
    if (tid == 0) out[1] = (t1 - t0) / N;
    // 3. dependent shuffle (64-bit = 2 SHFL) + add chain
    double s = f;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) s += __shfl_xor_sync(0xffffffffu, s, 1);
    t1 = clock64();
    if (tid == 0) out[2] = (t1 - t0) / N;
    // 4. dependent LDS chain (pointer chasing through shared memory)
    __shared__ int idx[256];
    idx[tid % 256] = (tid + 1) % 256;
    __syncthreads();
    int p = tid % 256;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) p = idx[p];
    t1 = clock64();
    if (tid == 0) out[3] = (t1 - t0) / N;
    // 5. __syncthreads round trip with all warps of the CTA
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) __syncthreads();
    t1 = clock64();
    if (tid == 0) out[4] = (t1 - t0) / N;
    // 6. FP64 division and sqrt chains
    double d = 3.0 + tid;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < 128; i++) d = 1.0 / d + 1.5;
    t1 = clock64();
    if (tid == 0) out[5] = (t1 - t0) / 128;
    double q = 3.0 + tid;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < 128; i++) q = sqrt(q) + 1.5;
    t1 = clock64();
    if (tid == 0) out[6] = (t1 - t0) / 128;
    double r = 3.0 + tid;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < 128; i++) r = rsqrt(r) + 1.5;
    t1 = clock64();
    if (tid == 0) out[7] = (t1 - t0) / 128;
    // 7. DMMA issue rate: 8 independent accumulators per warp
    double e[8][2];
    for (int j = 0; j < 8; j++) e[j][0] = e[j][1] = 0;
    t0 = clock64();
    for (int i = 0; i < 64; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) dmma(e[j][0], e[j][1], a, b);
    }
    t1 = clock64();
    if (tid == 0) out[8] = (t1 - t0) / 64;   // cycles per 8 independent DMMAs of one warp while all warps do the same
    double acc = c0 + c1 + f + s + p + d + q + r;
    for (int j = 0; j < 8; j++) acc += e[j][0] + e[j][1];
    sink[tid] = acc;
}
int main() {
    long long* out; double* sink;
    cudaMallocManaged(&out, 16 * sizeof(long long)); cudaMalloc(&sink, 1024 * sizeof(double));
    const char* names[] = {"DMMA.8x8x4 dependent", "DFMA dependent", "SHFL(64-bit)+DADD dependent", "LDS dependent", "__syncthreads", "1/x (FP64) + add", "sqrt (FP64) + add", "rsqrt (FP64) + add", "8 independent DMMA per warp"};
    for (int nt : {32, 128, 512}) {
        probe<<<1, nt>>>(out, sink, nt);
        cudaDeviceSynchronize();
        printf("threads=%d:", nt);
        for (int i = 0; i < 9; i++) printf("  %s=%lld", names[i], out[i]);
        printf("\n");
    }
    return 0;
}
