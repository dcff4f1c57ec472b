// sq_probe.cu — cycles per iteration of the trace-normalised squaring step used by pls_gram_kernel, with variants.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int VARIANT>
__global__ void __launch_bounds__(512, 1) sq(long long* out, double* sink, int Mp, int iters) {
    __shared__ double Sa[32 * 36], Sb[32 * 36], dgA[32], dgB[32];
    __shared__ unsigned char pta[16], ptb[16];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, gq = lane >> 2, qq = lane & 3;
    const int lds = Mp + 4, ntile = Mp / 8, npair = ntile * (ntile + 1) / 2;
    for (int i = tid; i < 32 * 36; i += 512) { Sa[i] = 0; Sb[i] = 0; }
    if (tid < 32) { dgA[tid] = 0; dgB[tid] = 0; }
    __syncthreads();
    for (int i = tid; i < Mp * Mp; i += 512) { const int r = i / Mp, c = i % Mp; const double v = 1.0 / (1.0 + abs(r - c)) + (r == c ? 1.0 + 0.01 * r : 0.0); Sa[r * lds + c] = v; if (r == c) dgA[r] = v; }
    if (tid == 0) { int p = 0; for (int a = 0; a < ntile; a++) for (int b = a; b < ntile; b++) { pta[p] = a; ptb[p] = b; p++; } }
    __syncthreads();
    const double* src = Sa; double* dst = Sb; const double* dgs = dgA; double* dgd = dgB;
    double tr = 0;
    for (int a = 0; a < Mp; a++) tr += dgs[a];
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        const int ex = ((__double2hiint(tr) >> 20) & 0x7ff) - 1023;
        const double sc = __hiloint2double((1023 - ex) << 20, 0), sc2 = sc * sc;
        if (VARIANT != 3) {
            for (int pidx = wid; pidx < npair; pidx += 16) {
                const int ta = pta[pidx], tb = ptb[pidx];
                const double* pa = src + (ta * 8 + gq) * lds + qq;
                const double* pb = src + (tb * 8 + gq) * lds + qq;
                double c[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
                int ks = 0;
                for (; ks + 16 <= Mp; ks += 16) {
#pragma unroll
                    for (int u = 0; u < 4; u++) dmma(c[u][0], c[u][1], pa[ks + 4 * u], pb[ks + 4 * u]);
                }
                if (ks < Mp) {
#pragma unroll
                    for (int u = 0; u < 4; u++) if (ks + 4 * u < Mp) dmma(c[u][0], c[u][1], pa[ks + 4 * u], pb[ks + 4 * u]);
                }
                const double c0 = ((c[0][0] + c[1][0]) + (c[2][0] + c[3][0])) * sc2, c1 = ((c[0][1] + c[1][1]) + (c[2][1] + c[3][1])) * sc2;
                const int r = ta * 8 + gq, cc = tb * 8 + 2 * qq;
                dst[r * lds + cc] = c0; dst[r * lds + cc + 1] = c1;
                if (VARIANT != 1) { if (ta != tb) { dst[cc * lds + r] = c0; dst[(cc + 1) * lds + r] = c1; } else if ((gq >> 1) == qq) dgd[r] = (gq & 1) ? c1 : c0; }
            }
        }
        __syncthreads();
        if (VARIANT != 2) {
            double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
            for (int a = 0; a < Mp; a += 8) { t0 += dgd[a] + dgd[a + 4]; t1 += dgd[a + 1] + dgd[a + 5]; t2 += dgd[a + 2] + dgd[a + 6]; t3 += dgd[a + 3] + dgd[a + 7]; }
            tr = (t0 + t1) + (t2 + t3);
        } else tr = tr * 0.5 + 1.0;
        src = dst; dst = (dst == Sa) ? Sb : Sa;
        { const double* tmp = dgs; dgs = dgd; dgd = (double*)tmp; }
        if (VARIANT == 0 && !(tr > 0.0)) break;
    }
    const long long t1 = clock64();
    if (tid == 0) out[0] = (t1 - t0) / iters;
    sink[tid] = tr + src[tid];
}
int main() {
    long long* out; double* sink;
    cudaMallocManaged(&out, 8 * sizeof(long long)); cudaMalloc(&sink, 1024 * sizeof(double));
    for (int Mp : {16, 32, 56}) {
        if (Mp > 32) continue;
        sq<0><<<1, 512>>>(out, sink, Mp, 200); cudaDeviceSynchronize(); printf("Mp=%d full=%lld", Mp, out[0]);
        sq<1><<<1, 512>>>(out, sink, Mp, 200); cudaDeviceSynchronize(); printf(" no-mirror/diag=%lld", out[0]);
        sq<2><<<1, 512>>>(out, sink, Mp, 200); cudaDeviceSynchronize(); printf(" no-diagsum=%lld", out[0]);
        sq<3><<<1, 512>>>(out, sink, Mp, 200); cudaDeviceSynchronize(); printf(" no-dmma(sync+diagsum only)=%lld\n", out[0]);
    }
    return 0;
}
