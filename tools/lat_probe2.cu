// lat_probe2.cu — sync and exchange latencies that decide the layout of the PLS component loop (cycles).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lat_probe2 tools/lat_probe2.cu && ./tools/lat_probe2
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__global__ void probe_cta(long long* out, double* sink) {
    __shared__ double sm[2048];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    sm[tid] = tid * 1e-3;
    __syncthreads();
    long long t0, t1;
    const int N = 256;
    double v = tid;
    // 1. named barrier among the first 2 / 4 / 8 warps
    for (int nw = 2, slot = 0; nw <= 8; nw <<= 1, slot++) {
        __syncthreads();
        t0 = clock64();
        if (wid < nw) {
#pragma unroll 8
            for (int i = 0; i < N; i++) asm volatile("bar.sync 1, %0;" ::"r"(nw * 32) : "memory");
        }
        t1 = clock64();
        if (tid == 0) out[slot] = (t1 - t0) / N;
    }
    // 2. same-warp STS -> __syncwarp -> LDS (other lane's value) dependent chain
    __syncthreads();
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; i++) { sm[tid] = v; __syncwarp(); v = sm[(tid ^ 1)] + 1.0; __syncwarp(); }
    t1 = clock64();
    if (tid == 0) out[3] = (t1 - t0) / N;
    // 3. cross-warp exchange through shared memory with a 2-warp named barrier: STS, bar, LDS
    __syncthreads();
    t0 = clock64();
    if (wid < 2) {
#pragma unroll 8
        for (int i = 0; i < N; i++) { sm[tid] = v; asm volatile("bar.sync 2, 64;" ::: "memory"); v = sm[tid ^ 32] + 1.0; asm volatile("bar.sync 2, 64;" ::: "memory"); }
    }
    t1 = clock64();
    if (tid == 0) out[4] = (t1 - t0) / N / 2;
    // 4. 5-round butterfly warp sum of a double (10 SHFL + 5 DADD), dependent
    __syncthreads();
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    t1 = clock64();
    if (tid == 0) out[5] = (t1 - t0) / N;
    // 5. DADD / DMUL dependent
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) v = v + 1.0;
    t1 = clock64();
    if (tid == 0) out[6] = (t1 - t0) / N;
    // 6. LDS.64 -> DFMA dependent (address from the value)
    int p = tid & 255;
    double acc = 0;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; i++) { const double x = sm[p]; acc = fma(x, 1.0000001, acc); p = (p + (int)(x > 1e30)) & 255; }
    t1 = clock64();
    if (tid == 0) out[7] = (t1 - t0) / N;
    // 7. __syncthreads with the whole CTA
    __syncthreads();
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; i++) __syncthreads();
    t1 = clock64();
    if (tid == 0) out[8] = (t1 - t0) / N;
    // 8. 32-bit shuffle dependent
    int iv = tid;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) iv += __shfl_xor_sync(0xffffffffu, iv, 1);
    t1 = clock64();
    if (tid == 0) out[9] = (t1 - t0) / N;
    sink[tid] = v + acc + p + iv;
}

// cluster probes: cluster.sync round trip; remote store -> cluster barrier -> local load
__global__ void probe_cluster(long long* out, double* sink) {
    __shared__ double sm[1024];
    cg::cluster_group cl = cg::this_cluster();
    const int tid = threadIdx.x;
    const unsigned rank = cl.block_rank(), nr = cl.num_blocks();
    sm[tid] = tid;
    cl.sync();
    long long t0, t1;
    const int N = 128;
    t0 = clock64();
    for (int i = 0; i < N; i++) cl.sync();
    t1 = clock64();
    if (tid == 0 && rank == 0) out[0] = (t1 - t0) / N;
    double v = tid;
    double* remote = cl.map_shared_rank(sm, (rank + 1) % nr);
    t0 = clock64();
    for (int i = 0; i < N; i++) { remote[tid] = v; cl.sync(); v = sm[tid] + 1.0; cl.sync(); }
    t1 = clock64();
    if (tid == 0 && rank == 0) out[1] = (t1 - t0) / N;
    // split arrive / wait with release-acquire, one barrier per exchange (double-buffered)
    t0 = clock64();
    for (int i = 0; i < N; i++) {
        double* dst = cl.map_shared_rank(sm + (i & 1) * 512, (rank + 1) % nr);
        dst[tid & 511] = v;
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        v = sm[(i & 1) * 512 + (tid & 511)] + 1.0;
    }
    t1 = clock64();
    if (tid == 0 && rank == 0) out[2] = (t1 - t0) / N;
    // remote load dependent chain
    const double* rsrc = cl.map_shared_rank(sm, (rank + 1) % nr);
    int p = tid & 255;
    t0 = clock64();
    for (int i = 0; i < N; i++) { const double x = rsrc[p]; p = (p + (int)(x > 1e30) + 1) & 255; }
    t1 = clock64();
    if (tid == 0 && rank == 0) out[3] = (t1 - t0) / N;
    cl.sync();
    sink[blockIdx.x * blockDim.x + tid] = v + p;
}

int main() {
    long long* out; double* sink;
    cudaMallocManaged(&out, 16 * sizeof(long long)); cudaMalloc(&sink, 16 * 1024 * sizeof(double));
    const char* names[] = {"bar.sync 2 warps", "bar.sync 4 warps", "bar.sync 8 warps", "STS+syncwarp+LDS+syncwarp", "STS+bar64+LDS (per bar)", "warp_sum(double)", "DADD dep", "LDS.64->DFMA dep", "__syncthreads", "SHFL32+IADD dep"};
    for (int nt : {256, 512}) {
        probe_cta<<<1, nt>>>(out, sink);
        cudaDeviceSynchronize();
        printf("cta threads=%d:", nt);
        for (int i = 0; i < 10; i++) printf("  %s=%lld", names[i], out[i]);
        printf("\n");
    }
    const char* cn[] = {"cluster.sync", "remote STS + 2 cluster.sync + LDS", "remote STS + arrive.release/wait.acquire + LDS", "remote LDS dep"};
    for (int cs : {2, 4, 8}) {
        for (int nt : {128, 512}) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cs); cfg.blockDim = dim3(nt);
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaError_t e = cudaLaunchKernelEx(&cfg, probe_cluster, out, sink);
            cudaDeviceSynchronize();
            printf("cluster=%d threads=%d (%s):", cs, nt, cudaGetErrorString(e));
            for (int i = 0; i < 4; i++) printf("  %s=%lld", cn[i], out[i]);
            printf("\n");
        }
    }
    return 0;
}
