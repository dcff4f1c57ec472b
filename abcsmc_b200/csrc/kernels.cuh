// kernels.cuh — internal launcher interface between api.cu and the kernel translation units.
#pragma once
#include "common.cuh"

// ---- moments.cu -------------------------------------------------------------------------------
size_t moments_ws_bytes(int64_t N, int K);
int launch_col_stats(abcb200_ctx* ctx, const double* X, int64_t ld, int64_t N, int K, double* stats, int* nchunk_out);
int launch_col_finalize(abcb200_ctx* ctx, const double* stats, int nchunk, int64_t N, int K, double* mean_out, double* sd_out,
                        double var_scale, double* var_out);
int launch_zscore(abcb200_ctx* ctx, const double* X, int64_t ld, int64_t N, int K, const double* stats, int nchunk,
                  const double* mean_in, const double* sd_in, double* Z, int64_t ldz, double* mean_out, double* sd_out,
                  const double* obs, double* obs_z);
int launch_gather_rows(abcb200_ctx* ctx, const double* src, int64_t ld, const uint64_t* idx, int64_t n, int P, double* out, int64_t ldo);
int launch_euclidean(abcb200_ctx* ctx, const double* S, int64_t ld, int64_t N, int K, const double* ref, double* out);

// ---- pls.cu -----------------------------------------------------------------------------------
// Device-resident factors of one PLS::Model (real parts; lib/PLS/include/PLS/pls.h:253).
// Convergence threshold of the trace-normalised squaring iteration (pls_defl.cu / pls_wide.cu / pls_gram.cu). The test
// tr(B_{j+1}) > (1 - d) (s_j tr B_j)^2 bounds the eigenvalue ratios of B_j: sum_{i>=2} l_i / l_1 < d / 2. It is evaluated one
// squaring late and the iterate that is used is B_{j+2}, whose ratios are the FOURTH power of those of B_j: d = 2e-5 makes it
// a projector to 1e-20 (d = 1e-9, used before, paid for about one more squaring per component and returned 1e-37); d = 2e-4 gives
// (1e-4)^4 = 1e-16, rounding level, for a third of a squaring less.
// Tried and dropped (round 2): stopping the squarings once the ratios of the iterate in hand are below ~1e-2 and letting ONE warp
// finish with ~7 products v <- B v. The squarings per component fell from 9.6 to 6.8 (C3) but a single warp's product costs 600-900
// cycles (no other warp hides its shared-memory and instruction latencies) against ~1000-1400 for a CTA-wide squaring that squares
// the ratio: the loop got 8-16 % slower (2.09 -> 2.25 / 2.43 ms at C3).
constexpr double PLS_EIG_DELTA = 2e-4;
struct PlsFactors {
    int K, M, A, method;
    int64_t n;           // training rows
    double *W, *P, *R;   // K x A, ld K
    double* Q;           // M x A, ld M
    double* T;           // n x A, ld ldt (nullable)
    int64_t ldt;
};
size_t pls_fit_ws_bytes(const abcb200_ctx* ctx, int64_t n, int K, int M, int method);
// Fits A components of kernel PLS on X (n x K), Y (n x M); factors' buffers are preallocated by the caller.
int pls_fit_dev(abcb200_ctx* ctx, const double* X, int64_t ldx, const double* Y, int64_t ldy, PlsFactors f);
// pls_gram.cu: Gram products + the persistent single-CTA component loop (fills W, P, R, Q; not T)
size_t pls_gram_ws_bytes(const abcb200_ctx* ctx, int64_t n, int K, int M);
int pls_fit_gram_dev(abcb200_ctx* ctx, const double* X, int64_t ldx, const double* Y, int64_t ldy, const PlsFactors& f);
// the component loop alone from XX = X^T X and XY = X^T Y (fills W, P, R, Q)
size_t pls_components_ws_bytes(int K, int M, int A);
int pls_components_dev(abcb200_ctx* ctx, const double* XX, const double* XY, const PlsFactors& f);
// pls_defl.cu: the same loop on the deflated Gram matrix, H and XY resident in shared memory (K <= 192)
bool pls_defl_fits(const abcb200_ctx* ctx, int K, int M);
size_t pls_defl_ws_bytes(int K, int A);
int pls_defl_dev(abcb200_ctx* ctx, const double* XX, const double* XY, const PlsFactors& f, long long* prof);
// Model::cv_LOO as batched on-chip refits from down-dated Gram matrices (one persistent CTA per SM walks the held-out rows)
int pls_ur_dev(abcb200_ctx* ctx, const PlsFactors& f, double* U);
int pls_ur_block_dev(abcb200_ctx* ctx, const PlsFactors& f, double* U, int a_begin, int a_end);
// chunked fit: components [c0, c1) per launch, the loop state handed over through `state` (pls_defl_state_doubles)
size_t pls_defl_state_doubles(int K, int M);
int pls_defl_chunk_dev(abcb200_ctx* ctx, const double* XX, const double* XY, const PlsFactors& f, int c0, int c1, double* state, long long* prof);
// pls_wide.cu: the component loop for wide predictor sets (H and XY in L2, three launches per component)
bool pls_wide_fits(const abcb200_ctx* ctx, int K, int M);
size_t pls_wide_ws_bytes(int K, int M, int A);
int pls_wide_dev(abcb200_ctx* ctx, const double* XX, const double* XY, const PlsFactors& f);
// the same loop in blocks of components (pipelined ranking): state lives in global memory, so a block is just its launches
struct WideJob { int K, M, A; const double* XY0; double *H, *XYb[2], *wb[2], *pb[2], *qb[2], *S0, *W, *P, *Q; };
int pls_wide_begin(abcb200_ctx* ctx, const double* XX, const double* XY, const PlsFactors& f, WideJob* job);
int pls_wide_block(abcb200_ctx* ctx, const WideJob* job, int c0, int c1);
// sample.cu: next-set proposal sampling (weighted draw of predictive-prior rows + truncated normal noise)
size_t sample_ws_bytes(int64_t n_pp);
int sample_predictive_priors_core(abcb200_ctx* ctx, uint64_t seed, int64_t num_samples, const double* weights, const double* theta,
                                  int64_t ld, int64_t n_pp, int P, const double* dv, const double* lo, const double* hi,
                                  const int32_t* integral, const double* prior_mean, int max_attempts, double* out, int64_t ld_out,
                                  uint64_t* parent, unsigned long long* fallbacks);
size_t mvn_setup_ws_bytes(int64_t n_pp, int P);
int setup_mvn_sampler_core(abcb200_ctx* ctx, const double* theta, int64_t ld, int64_t n_pp, int P, double* L, int* flag);
int sample_mvn_predictive_priors_core(abcb200_ctx* ctx, uint64_t seed, int64_t num_samples, const double* weights, const double* theta,
                                      int64_t ld, int64_t n_pp, int P, const double* L, const double* lo, const double* hi,
                                      const int32_t* integral, int max_attempts, double* out, int64_t ld_out, uint64_t* parent,
                                      unsigned long long* failures);
bool pls_loo_fits(const abcb200_ctx* ctx, int K, int M);
int pls_loo_dev(abcb200_ctx* ctx, const double* X, int64_t ldx, const double* Y, int64_t ldy, int64_t N, int K, int M, int A,
                const double* XX, const double* XY, double* cube);
// out (n x ncols, ldo) = X (n x K) * B[:, :ncols] (K x ncols, ldb)
// (rows >= split are stored `gap` rows further down: see xb_kernel)
int launch_xb(abcb200_ctx* ctx, const double* X, int64_t ldx, int64_t n, int K, const double* B, int64_t ldb, int ncols,
              double* out, int64_t ldo, int64_t split = INT64_MAX, int64_t gap = 0);
// dist[i] = || X[i,:] * B[:, :ncols] - ref_scores ||_2   (projection fused with ABC::euclidean)
int launch_project_dist(abcb200_ctx* ctx, const double* X, int64_t ldx, int64_t n, int K, const double* B, int64_t ldb, int ncols,
                        const double* ref_scores, double* dist);
// out[a] = sum_k v[k] * B[k, a]
int launch_vec_times_mat(abcb200_ctx* ctx, const double* v, int K, const double* B, int64_t ldb, int ncols, double* out);
// C (Ka x Kb, ld Ka) = A^T B over n rows (DMMA, split over row chunks, deterministic reduction)
size_t atb_ws_bytes(const abcb200_ctx* ctx, int64_t n, int Ka, int Kb);
int launch_atb(abcb200_ctx* ctx, const double* A, int64_t lda, int Ka, const double* B, int64_t ldb, int Kb, int64_t n, double* C);
// gram.cu: XX (K x K, both triangles) = X^T X and XY (K x M) = X^T Y in one pass (TMA ring + DMMA, deterministic)
size_t gram_ws_bytes(const abcb200_ctx* ctx, int64_t n, int K, int M);
int launch_gram(abcb200_ctx* ctx, const double* X, int64_t ldx, int K, const double* Y, int64_t ldy, int M, int64_t n, double* XX, double* XY);
// C (K x M) = R[:, :comp] Q[:, :comp]^T  (Model::coefficients)
int launch_coefficients(abcb200_ctx* ctx, const double* R, const double* Q, int K, int M, int comp, double* C);

// ---- holdout.cu -------------------------------------------------------------------------------
size_t holdout_ws_bytes(const abcb200_ctx* ctx, int64_t n_te, int K, int M, int A, bool own_scores = true);
// The pieces of holdout_select_dev, for callers that feed the validation block by block (the pipelined ranking, api.cu)
struct HoldoutJob {
    int64_t n_te, ldt, ldn, ldy, rows_per_split;
    int K, M, A, nchk, nblk, ycta, exact_cap, ngroup, nsplit;
    double* T;                 // hold-out scores n_te x A (ld ldt): own buffer or the caller's
    const double *Yte, *Q;
    double *partial, *press, *chk, *Eref;
    int *ref, *decided, *result, *status, *work1, *work2, *summ;
    void* info;
    unsigned int *ghist, *ticket;
    uint32_t* s2hist;
    uint32_t* xs;              // level-2 fine-bin ranks kept for the exact level (holdout.cu: XS_BYTES)
};
int holdout_begin(abcb200_ctx* ctx, const double* Yte, int64_t ldy, int64_t n_te, const PlsFactors& f, const double* T_ext, int64_t ldt_ext,
                  double* press_dev, HoldoutJob* job);
int holdout_scores(abcb200_ctx* ctx, const HoldoutJob* job, const double* Zte, int64_t ldx, const double* R);
int holdout_press_block(abcb200_ctx* ctx, const HoldoutJob* job, int c_begin, int c_end);
int holdout_press_finalize(abcb200_ctx* ctx, const HoldoutJob* job);
int holdout_select_finish(abcb200_ctx* ctx, const HoldoutJob* job, double alpha, int32_t* ncomp_host);
// Streams cv_NEW_DATA + validation(RESS) + optimal_num_components. press_dev: M x A col-major (device, nullable);
// ncomp_host: M entries (host). Zte/Yte are the standardised hold-out rows.
int holdout_select_dev(abcb200_ctx* ctx, const double* Zte, int64_t ldx, const double* Yte, int64_t ldy, int64_t n_te,
                       const PlsFactors& f, double alpha, double* press_dev, int32_t* ncomp_host);
// Single signed-rank test on the device (PLS::wilcoxon)
size_t wilcoxon_ws_bytes(int64_t n);
int wilcoxon_dev(abcb200_ctx* ctx, const double* e1, const double* e2, int64_t n, double* p_host);

// ---- loo.cu -----------------------------------------------------------------------------------
// Residual cube layout (PLS::Residual): cube[(y * A + c) * n + i] = error of response y, row i, with c + 1 components.
size_t cube_select_ws_bytes(int64_t n, int M, int A);
int cube_select_dev(abcb200_ctx* ctx, const double* cube, int64_t n, int M, int A, int out_type, double alpha, double* press_host,
                    int32_t* ncomp_host);
// cube rows [row0, row0 + n) from scores T (n x A) and responses Y (n x M): e_c = e_{c-1} - t_c q_c; add accumulates
int launch_cube_from_scores(abcb200_ctx* ctx, const double* T, int64_t ldt, const double* Y, int64_t ldy, const double* Q, int64_t n, int M, int A,
                            double* cube, int64_t cube_n, int64_t row0, bool add);
// XXo = XX - x x^T, XYo = XY - x y^T with x = X[row, :], y = Y[row, :] (leave-one-out Gram matrices)
int launch_gram_downdate(abcb200_ctx* ctx, const double* XX, const double* XY, const double* X, int64_t ldx, const double* Y, int64_t ldy,
                         int64_t row, int K, int M, double* XXo, double* XYo);

// ---- sort.cu ----------------------------------------------------------------------------------
// Stable LSD radix sort of n_seg equal-length segments of 64-bit keys (optional 32-bit payload).
// keys/vals are sorted in place using alt buffers of the same size; hist is scratch from radix_ws_bytes.
size_t radix_hist_bytes(int64_t seg_len, int n_seg);
int radix_sort_segments(abcb200_ctx* ctx, uint64_t* keys, uint64_t* keys_alt, uint32_t* vals, uint32_t* vals_alt,
                        int64_t seg_len, int n_seg, uint32_t* hist, const int* seg_valid /*device, nullable*/);
size_t order_ws_bytes(int64_t n);
// order_out[0..top_n) = indices of the top_n smallest v (ascending, ties by index). Returns ABCB200_ENAN on NaN input.
int order_dev(abcb200_ctx* ctx, const double* v, int64_t n, int64_t top_n, uint64_t* order_out);

// ---- weights.cu -------------------------------------------------------------------------------
size_t weights_ws_bytes(const abcb200_ctx* ctx, int64_t n_new, int64_t n_old, int P);
int weights_unnorm_dev(abcb200_ctx* ctx, const double* numer, const double* th_new, int64_t ld_new, int64_t n_new,
                       const double* th_old, int64_t ld_old, int64_t n_old, const double* w_old, const double* dv_old, int P,
                       int algo, double* w_out, double* sumsq_out);
// the two phases of weights_unnorm_dev, for callers that exchange the conditioning maximum between them (sharded.cu)
struct WeightsJob {
    const double *th_new, *th_old, *w_old;
    int64_t ld_new, n_new, ld_old, n_old;
    int P, algo, nfin;
    double *scale, *centre, *scal, *Apk, *Bpk, *ss_part;
    int *poison, *nanflag;
};
int weights_pack(abcb200_ctx* ctx, const double* th_new, int64_t ld_new, int64_t n_new, const double* th_old, int64_t ld_old,
                 int64_t n_old, const double* w_old, const double* dv_old, int P, int algo, WeightsJob* job);
int weights_eval(abcb200_ctx* ctx, const WeightsJob* job, const double* numer, double* w_out, double* sumsq_out);
int launch_scale_weights(abcb200_ctx* ctx, double* w, int64_t n, const double* sumsq);
int launch_fill(abcb200_ctx* ctx, double* p, int64_t n, double v);

// ---- api.cu: the ranking on device-resident inputs (used by chain.cu) --------------------------------------------------
// Exact distance ties placed as libstdc++'s std::sort leaves them (abcb200_set_tie_order 1; api.cu). Host arrays: dist (N), order (top_n,
// the device order, rewritten in place when ties reach the output). The _device form brings both to the host, and the order back if it changed.
bool tie_order_stdsort(const double* dist, int64_t N, int64_t top_n, uint64_t* order);
int tie_order_stdsort_device(abcb200_ctx* ctx, const double* d_dist, int64_t N, int64_t top_n, uint64_t* d_order);
size_t rank_ws_bytes(const abcb200_ctx* ctx, int64_t N, int K, int P, double f, int method, bool simple);
int rank_shape_check(abcb200_ctx* ctx, int64_t N, int K, int P, double f, int method, bool simple);
// Column blocks of the inputs still on their way to the device (copied on ctx->copy_stream): block b of the metrics is columns
// [col0[b], col0[b + 1]) and is complete when ev[b] fires; ev_par covers the target and all the parameter columns.
struct Arrival { int nblk; int col0[11]; cudaEvent_t ev[10]; cudaEvent_t ev_par; };
int stage_inputs(abcb200_ctx* ctx, double* d_met, double* d_par, double* d_target, int64_t ldd, const double* met, int64_t ld_met, const double* par,
                 int64_t ld_par, const double* target, int64_t N, int K, int P, Arrival* arr);
int rank_on_device(abcb200_ctx* ctx, const double* met, int64_t ld_met, const double* par, int64_t ld_par, int64_t N, int K, int P, const double* target,
                   double f, int method, int64_t top_n, uint64_t* order_out, double* dist_out, int* n_comp_used_host, int32_t* n_comp_host, bool simple,
                   const Arrival* arr = nullptr);
