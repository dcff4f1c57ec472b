// moments.cu — S1: column moments and standardisation, plus the small
// column reductions of row a11 (doubled variance) and ABC::euclidean (a9) for the free-function API.
//
// Reference: PLS::SST / colwise_stdev / z_scores / colwise_z_scores, lib/PLS/src/pls.cpp:69-111;
//            ABC::calculate_doubled_variance, src/AbcUtil.cpp:528-537; ABC::euclidean, :320-324.
// Layout: column-major N x K with leading dimension ld, so every column is one contiguous HBM stream.
// Bound: HBM. Algorithmic bytes: 8*N*K read for the moments; 8*N*K read + 8*N*K written for z.
#include "kernels.cuh"

namespace {

constexpr int MOM_THREADS = 256;
constexpr int MOM_VPT = 8;                       // values per thread held in registers
constexpr int MOM_CHUNK = MOM_THREADS * MOM_VPT; // rows per CTA

// One CTA per (row chunk, column): exact two-pass mean / sum of squared deviations of the chunk from registers.
// stats[(col * nchunk + chunk) * 2 + {0,1}] = {chunk mean, chunk M2}
__global__ void __launch_bounds__(MOM_THREADS) col_chunk_stats_kernel(const double* __restrict__ X, int64_t ld, int64_t N,
                                                                      int nchunk, double* __restrict__ stats) {
    __shared__ double red[32];
    const int chunk = blockIdx.x, col = blockIdx.y;
    const int64_t r0 = (int64_t)chunk * MOM_CHUNK;
    const int64_t nrow = min((int64_t)MOM_CHUNK, N - r0);
    const double* x = X + (int64_t)col * ld + r0;
    double v[MOM_VPT];
    double s = 0;
#pragma unroll
    for (int j = 0; j < MOM_VPT; j++) {
        const int64_t i = (int64_t)j * MOM_THREADS + threadIdx.x;
        v[j] = (i < nrow) ? x[i] : 0.0;
        s += v[j];
    }
    s = block_sum(s, red);
    const double mean = s / (double)nrow;
    double m2 = 0;
#pragma unroll
    for (int j = 0; j < MOM_VPT; j++) {
        const int64_t i = (int64_t)j * MOM_THREADS + threadIdx.x;
        const double d = v[j] - mean;
        m2 += (i < nrow) ? d * d : 0.0;
    }
    m2 = block_sum(m2, red);
    if (threadIdx.x == 0) {
        double* o = stats + ((int64_t)col * nchunk + chunk) * 2;
        o[0] = mean;
        o[1] = m2;
    }
}

// Merge chunk statistics of one column (exact pooled formulas, fixed order): one warp.
__device__ __forceinline__ void merge_chunks(const double* __restrict__ st, int nchunk, int64_t N, double& mean, double& var) {
    const int lane = threadIdx.x & 31;
    double s1 = 0;
    for (int c = lane; c < nchunk; c += 32) {
        const double nc = (double)min((int64_t)MOM_CHUNK, N - (int64_t)c * MOM_CHUNK);
        s1 += nc * st[2 * c];
    }
    s1 = warp_sum(s1);
    mean = s1 / (double)N;
    double m2 = 0;
    for (int c = lane; c < nchunk; c += 32) {
        const double nc = (double)min((int64_t)MOM_CHUNK, N - (int64_t)c * MOM_CHUNK);
        const double d = st[2 * c] - mean;
        m2 += st[2 * c + 1] + nc * d * d;
    }
    m2 = warp_sum(m2);
    if (N < 2) m2 = 0.0;                       // PLS::SST returns zeros for N < 2 (pls.cpp:71)
    var = m2 / ((double)N - 1.0);              // pls.cpp:82 (sd = sqrt(var)); N == 1 -> 0/0 = NaN as in the reference
}

__global__ void col_finalize_kernel(const double* __restrict__ stats, int nchunk, int64_t N, int K,
                                    double* __restrict__ mean_out, double* __restrict__ sd_out, double var_scale,
                                    double* __restrict__ var_out) {
    const int col = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (col >= K) return;
    double mean, var;
    merge_chunks(stats + (int64_t)col * nchunk * 2, nchunk, N, mean, var);
    if ((threadIdx.x & 31) == 0) {
        if (mean_out) mean_out[col] = mean;
        if (sd_out) sd_out[col] = sqrt(var);
        if (var_out) var_out[col] = (N > 1) ? var_scale * var : 0.0;   // RunningStat::Variance: n<2 -> 0
    }
}

// z = (x - mean) / sd with the UNGUARDED sd (pls.cpp:103). When stats != nullptr the column's mean/sd are
// merged from chunk statistics by warp 0 of every CTA (saves a launch); CTAs with blockIdx.x == 0 publish them
// and the standardised observation (pls.cpp:89-91).
__global__ void __launch_bounds__(MOM_THREADS) zscore_kernel(const double* __restrict__ X, int64_t ld, int64_t N, int K,
                                                     const double* __restrict__ stats, int nchunk,
                                                     const double* __restrict__ mean_in, const double* __restrict__ sd_in,
                                                     double* __restrict__ Z, int64_t ldz, double* __restrict__ mean_out,
                                                     double* __restrict__ sd_out, const double* __restrict__ obs,
                                                     double* __restrict__ obs_z) {
    __shared__ double sh[2];
    const int col = blockIdx.y;
    // the CTA's MOM_CHUNK rows are requested first: their HBM latency overlaps the merge of the chunk statistics below
    const int64_t r0 = (int64_t)blockIdx.x * MOM_CHUNK;
    const int64_t nrow = min((int64_t)MOM_CHUNK, N - r0);
    double v[MOM_VPT];
    {
        const double* xin = X + (int64_t)col * ld + r0;
#pragma unroll
        for (int j = 0; j < MOM_VPT; j++) {
            const int64_t i = (int64_t)j * MOM_THREADS + threadIdx.x;
            v[j] = (i < nrow) ? xin[i] : 0.0;
        }
    }
    if (stats && nchunk > 64) {
        // long sets: the whole CTA merges (two block sums) instead of one warp walking nchunk / 32 dependent trips
        __shared__ double red[32];
        const double* st = stats + (int64_t)col * nchunk * 2;
        double s1 = 0;
        for (int c = threadIdx.x; c < nchunk; c += MOM_THREADS) s1 += (double)min((int64_t)MOM_CHUNK, N - (int64_t)c * MOM_CHUNK) * st[2 * c];
        const double mean = block_sum(s1, red) / (double)N;
        double m2 = 0;
        for (int c = threadIdx.x; c < nchunk; c += MOM_THREADS) {
            const double d = st[2 * c] - mean;
            m2 += st[2 * c + 1] + (double)min((int64_t)MOM_CHUNK, N - (int64_t)c * MOM_CHUNK) * d * d;
        }
        m2 = block_sum(m2, red);
        if (threadIdx.x == 0) { sh[0] = mean; sh[1] = sqrt(m2 / ((double)N - 1.0)); }
    } else if (stats) {
        if (threadIdx.x < 32) {
            double mean, var;
            merge_chunks(stats + (int64_t)col * nchunk * 2, nchunk, N, mean, var);
            if (threadIdx.x == 0) { sh[0] = mean; sh[1] = sqrt(var); }
        }
    } else if (threadIdx.x == 0) {
        sh[0] = mean_in[col]; sh[1] = sd_in[col];
    }
    __syncthreads();
    const double mean = sh[0], sd = sh[1];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (mean_out) mean_out[col] = mean;
        if (sd_out) sd_out[col] = sd;
        if (obs_z) obs_z[col] = (obs[col] - mean) / sd;
    }
    double* z = Z + (int64_t)col * ldz + r0;
#pragma unroll
    for (int j = 0; j < MOM_VPT; j++) {
        const int64_t i = (int64_t)j * MOM_THREADS + threadIdx.x;
        if (i < nrow) z[i] = (v[j] - mean) / sd;
    }
}

// rows gathered by index: out[i, p] = src[idx[i], p]
__global__ void gather_rows_kernel(const double* __restrict__ src, int64_t ld, const uint64_t* __restrict__ idx, int64_t n,
                                   int P, double* __restrict__ out, int64_t ldo) {
    const int p = blockIdx.y;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[(int64_t)p * ldo + i] = src[(int64_t)p * ld + (int64_t)min(idx[i], (uint64_t)(ld - 1))];   // an index past the column never reads outside the matrix
}

// ABC::euclidean (src/AbcUtil.cpp:320-324): d_i = sqrt(sum_k (S[i,k] - ref[k])^2), thread per row, coalesced per column
__global__ void euclidean_kernel(const double* __restrict__ S, int64_t ld, int64_t N, int K, const double* __restrict__ ref,
                                 double* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        double acc = 0;
        for (int k = 0; k < K; k++) {
            const double d = S[(int64_t)k * ld + i] - ref[k];
            acc = fma(d, d, acc);
        }
        out[i] = sqrt(acc);
    }
}

}  // namespace

size_t moments_ws_bytes(int64_t N, int K) {
    const int nchunk = (int)((N + MOM_CHUNK - 1) / MOM_CHUNK);
    return align_up((size_t)K * nchunk * 2 * sizeof(double), 256) + 256;
}

int launch_col_stats(abcb200_ctx* ctx, const double* X, int64_t ld, int64_t N, int K, double* stats, int* nchunk_out) {
    const int nchunk = (int)((N + MOM_CHUNK - 1) / MOM_CHUNK);
    *nchunk_out = nchunk;
    LAUNCH(ctx, col_chunk_stats_kernel, dim3(nchunk, K), MOM_THREADS, 0, X, ld, N, nchunk, stats);
    return ABCB200_OK;
}

int launch_col_finalize(abcb200_ctx* ctx, const double* stats, int nchunk, int64_t N, int K, double* mean_out, double* sd_out,
                        double var_scale, double* var_out) {
    LAUNCH(ctx, col_finalize_kernel, (K + 3) / 4, 128, 0, stats, nchunk, N, K, mean_out, sd_out, var_scale, var_out);
    return ABCB200_OK;
}

int launch_zscore(abcb200_ctx* ctx, const double* X, int64_t ld, int64_t N, int K, const double* stats, int nchunk,
                  const double* mean_in, const double* sd_in, double* Z, int64_t ldz, double* mean_out, double* sd_out,
                  const double* obs, double* obs_z) {
    const int gx = (int)max((int64_t)1, (N + MOM_CHUNK - 1) / MOM_CHUNK);
    LAUNCH(ctx, zscore_kernel, dim3(gx, K), MOM_THREADS, 0, X, ld, N, K, stats, nchunk, mean_in, sd_in, Z, ldz, mean_out, sd_out, obs, obs_z);
    return ABCB200_OK;
}

int launch_gather_rows(abcb200_ctx* ctx, const double* src, int64_t ld, const uint64_t* idx, int64_t n, int P, double* out, int64_t ldo) {
    int gx = (int)max((int64_t)1, min((n + 255) / 256, (int64_t)1024));
    LAUNCH(ctx, gather_rows_kernel, dim3(gx, P), 256, 0, src, ld, idx, n, P, out, ldo);
    return ABCB200_OK;
}

int launch_euclidean(abcb200_ctx* ctx, const double* S, int64_t ld, int64_t N, int K, const double* ref, double* out) {
    int gx = (int)max((int64_t)1, min((N + 255) / 256, (int64_t)(16 * ctx->sm_count)));
    LAUNCH(ctx, euclidean_kernel, gx, 256, 0, S, ld, N, K, ref, out);
    return ABCB200_OK;
}
