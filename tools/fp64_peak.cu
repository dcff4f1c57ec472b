// fp64_peak.cu — measures the FP64 roofline denominators on the box (B200, sm_100a):
//   (1) vector DFMA peak, (2) DMMA (mma.sync.m8n8k4.f64) peak, (3) both interleaved,
//   (4) libdevice exp() vs the hand-rolled exp(-q) used by the weight-update kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
    double c0 = threadIdx.x, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
    for (int i = 0; i < iters; i++) {
        c0 = fma(c0, a, b); c1 = fma(c1, a, b); c2 = fma(c2, a, b); c3 = fma(c3, a, b);
        c4 = fma(c4, a, b); c5 = fma(c5, a, b); c6 = fma(c6, a, b); c7 = fma(c7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double a, double b) {
    double c[8][2];
    for (int j = 0; j < 8; j++) { c[j][0] = threadIdx.x + j; c[j][1] = j; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) dmma(c[j][0], c[j][1], a, b);
    }
    double s = 0; for (int j = 0; j < 8; j++) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) mixed_kernel(double* out, int iters, double a, double b) {
    double c[4][2];
    for (int j = 0; j < 4; j++) { c[j][0] = threadIdx.x + j; c[j][1] = j; }
    double f0 = threadIdx.x, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3, f4 = f0 + 4, f5 = f0 + 5, f6 = f0 + 6, f7 = f0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 4; j++) dmma(c[j][0], c[j][1], a, b);
        f0 = fma(f0, a, b); f1 = fma(f1, a, b); f2 = fma(f2, a, b); f3 = fma(f3, a, b);
        f4 = fma(f4, a, b); f5 = fma(f5, a, b); f6 = fma(f6, a, b); f7 = fma(f7, a, b);
    }
    double s = f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7; for (int j = 0; j < 4; j++) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// hand-rolled exp(-q), q >= 0: n = rint(-q*log2e), r = -q - n*ln2 (hi/lo), degree-11 Taylor/Horner, scale by 2^n
__device__ __forceinline__ double exp_neg(double q) {
    const double x = -q;
    const double L2E = 1.4426950408889634074, LN2HI = 6.93147180369123816490e-01, LN2LO = 1.90821492927058770002e-10;
    double t = fma(x, L2E, 6755399441055744.0);
    int n = __double2loint(t);
    double fn = t - 6755399441055744.0;
    double r = fma(fn, -LN2HI, x); r = fma(fn, -LN2LO, r);
    double p = 2.505210838544172e-08;   // 1/11!
    p = fma(p, r, 2.755731922398589e-07); p = fma(p, r, 2.755731922398589e-06); p = fma(p, r, 2.480158730158730e-05);
    p = fma(p, r, 1.984126984126984e-04); p = fma(p, r, 1.388888888888889e-03); p = fma(p, r, 8.333333333333333e-03);
    p = fma(p, r, 4.166666666666666e-02); p = fma(p, r, 1.666666666666667e-01); p = fma(p, r, 0.5);
    p = fma(p, r, 1.0); p = fma(p, r, 1.0);
    int hi = __double2hiint(p) + (n << 20);
    double res = __hiloint2double(hi, __double2loint(p));
    return (n < -1021) ? 0.0 : res;
}

__global__ void __launch_bounds__(256) exp_lib_kernel(double* out, int iters, double q0) {
    double q = q0 + threadIdx.x * 1e-3, s = 0;
    for (int i = 0; i < iters; i++) { s += exp(-q); q += 1e-3; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) exp_own_kernel(double* out, int iters, double q0) {
    double q = q0 + threadIdx.x * 1e-3, s = 0;
    for (int i = 0; i < iters; i++) { s += exp_neg(q); q += 1e-3; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void exp_check_kernel(double* maxrel, int n) {
    double worst = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double q = 745.0 * i / n;
        double a = exp(-q), b = exp_neg(q);
        if (a > 1e-300) { double rel = fabs(a - b) / a; if (rel > worst) worst = rel; }
    }
    atomicMax((unsigned long long*)maxrel, (unsigned long long)__double_as_longlong(worst));
}

template <typename F> float timeit(F f, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) { CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", prop.name, sms, prop.clockRate);
    int blocks = sms * 8, threads = 256, iters = 20000;
    double* out; CK(cudaMalloc(&out, sizeof(double) * blocks * threads));
    float ms;
    ms = timeit([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
    double dfma_tf = 2.0 * 8 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    printf(" \"dfma_tflops\": %.3f, \"dfma_ms\": %.3f,\n", dfma_tf, ms);
    ms = timeit([&] { dmma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
    double dmma_tf = 2.0 * 8 * 256.0 * iters * (double)blocks * (threads / 32) / (ms * 1e-3) / 1e12;
    printf(" \"dmma_tflops\": %.3f, \"dmma_ms\": %.3f,\n", dmma_tf, ms);
    ms = timeit([&] { mixed_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
    double mixed_tf = (2.0 * 4 * 256.0 * iters * (double)blocks * (threads / 32) + 2.0 * 8 * iters * (double)blocks * threads) / (ms * 1e-3) / 1e12;
    printf(" \"mixed_dmma4_dfma8_tflops\": %.3f, \"mixed_ms\": %.3f,\n", mixed_tf, ms);
    int eiters = 4000;
    ms = timeit([&] { exp_lib_kernel<<<blocks, threads>>>(out, eiters, 0.5); }, 5);
    printf(" \"exp_lib_gexp_s\": %.2f,\n", (double)eiters * blocks * threads / (ms * 1e-3) / 1e9);
    ms = timeit([&] { exp_own_kernel<<<blocks, threads>>>(out, eiters, 0.5); }, 5);
    printf(" \"exp_own_gexp_s\": %.2f,\n", (double)eiters * blocks * threads / (ms * 1e-3) / 1e9);
    double* mr; CK(cudaMalloc(&mr, 8)); CK(cudaMemset(mr, 0, 8));
    exp_check_kernel<<<1, 1024>>>(mr, 4000000); double h; CK(cudaMemcpy(&h, mr, 8, cudaMemcpyDeviceToHost));
    printf(" \"exp_own_max_rel_err\": %.3e,\n", h);
    // HBM copy sanity (read+write bytes)
    size_t nb = (size_t)2 << 30; char *a, *b; CK(cudaMalloc(&a, nb)); CK(cudaMalloc(&b, nb)); CK(cudaMemset(a, 1, nb));
    ms = timeit([&] { cudaMemcpyAsync(b, a, nb, cudaMemcpyDeviceToDevice); }, 5);
    printf(" \"d2d_copy_gbs\": %.1f}\n", 2.0 * nb / (ms * 1e-3) / 1e9);
    return 0;
}
