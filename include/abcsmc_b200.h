/* abcsmc_b200.h — C ABI of the B200-native (sm_100a) implementation of AbcSmc's per-set
 * post-simulation hot path (PLS ranking, top-N selection, doubled variance, SMC weight update).
 *
 * The reference (tjhladish/AbcSmc) has no FFI layer for this path; its boundary is the C++ API in
 * include/AbcSmc/AbcUtil.h:146-172 and lib/PLS/include/PLS/pls.h:58-266. Each entry point below names
 * the reference function it replaces. A header-only C++ adapter (abcsmc_b200/host/abc_b200_adapter.hpp)
 * re-creates the reference signatures on top of this ABI; INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - All matrices are FP64, column-major (Eigen::MatrixXd layout), with an explicit leading dimension
 *    `ld` (in elements, >= rows). Index outputs are uint64_t (the reference's size_t).
 *  - Functions without a `_dev` suffix take HOST pointers and perform the H2D/D2H copies themselves;
 *    `_dev` variants take DEVICE pointers valid on the context's device and run on the context's stream
 *    (asynchronously unless they return a host scalar).
 *  - Return value: 0 on success, a negative ABCB200_E* code otherwise (the reference asserts/exits; this
 *    library never exits). abcb200_last_error() returns a message for the last failure on a context.
 *  - A context is not thread-safe; use one per host thread. There is no CPU fallback: creating a context
 *    without a CUDA device fails with ABCB200_ENODEV.
 */
#ifndef ABCSMC_B200_H
#define ABCSMC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABCB200_OK 0
#define ABCB200_EINVAL (-1)   /* bad shape/argument (the reference would assert) */
#define ABCB200_ENODEV (-2)   /* no usable CUDA device */
#define ABCB200_ECUDA (-3)    /* CUDA runtime failure, see abcb200_last_error */
#define ABCB200_ENOMEM (-4)
#define ABCB200_ENAN (-5)     /* NaN in distances: ordering undefined in the reference (std::sort UB) */

/* PLS::METHOD, lib/PLS/include/PLS/pls.h:131 */
#define ABCB200_KERNEL_TYPE1 0
#define ABCB200_KERNEL_TYPE2 1
/* Same results as KERNEL_TYPE1 evaluated literally as pls.cpp:418-421 does: X is streamed once per component
 * (t = X r, p = X^T t / t^T t). TYPE1 and TYPE2 both run from the Gram matrices X^T X, X^T Y (one read of X);
 * use this variant when X^T X is too ill-conditioned for the Gram form (tt = r^T XX r cancels). */
#define ABCB200_KERNEL_TYPE1_STREAM 2
/* PLS::VALIDATION_OUTPUT, lib/PLS/include/PLS/pls.h:143 */
#define ABCB200_RESS 0
#define ABCB200_MSE 1

typedef struct abcb200_ctx abcb200_ctx;
typedef struct abcb200_pls abcb200_pls;   /* PLS::Model, lib/PLS/include/PLS/pls.h:184-266 */

/* ---- context ------------------------------------------------------------------------------- */
int abcb200_create(int device, abcb200_ctx** out);
int abcb200_destroy(abcb200_ctx* ctx);
/* Run on an externally owned cudaStream_t (e.g. torch's current stream; NULL = CUDA's legacy default stream).
 * ABCB200_OWN_STREAM restores the context's own non-blocking stream. */
#define ABCB200_OWN_STREAM ((void*)(intptr_t)-1)
int abcb200_set_stream(abcb200_ctx* ctx, void* cuda_stream);
int abcb200_synchronize(abcb200_ctx* ctx);
const char* abcb200_last_error(abcb200_ctx* ctx);
/* CUDA-event instrumentation, OFF by default: stage_on switches the per-stage brackets (abcb200_stage_ms), bit k of kernel_mask
 * the bracket of hot kernel k (abcb200_kernel_ms). Every record is a stream operation between launches: at the small shapes
 * (35 launches in 0.6 ms) twenty of them cost 14 % of the step, so a timed run should switch on only what it reports.
 * ABCB200_TIMERS=1|2|3 (stages | all kernels | both) overrides the default at context creation (debugging). */
int abcb200_set_timers(abcb200_ctx* ctx, int stage_on, uint32_t kernel_mask);
/* Placement of EXACT distance ties in every order this context returns (abcb200_rank_* and the chained entry points).
 * 0 (default): ascending particle index. 1: where libstdc++'s std::sort leaves them in PLS::ordered (lib/PLS/include/PLS/pls.h:58-69:
 * an index sort with a strict <, not stable — the placement is a property of the introsort run over all N indices, see also
 * lib/ranker.h:47-53). Mode 1 brings the distances (8 N bytes) to the host after the ranking; when no two of the returned distances
 * are equal and none beyond the cut equals the last one, the device order already IS what std::sort returns (the common case:
 * continuous metrics) and nothing else happens; otherwise the same std::sort runs on the host over the GPU-computed distances and
 * its first top_n entries are returned. abcb200_stat(ctx, 8) counts the rankings that needed it. */
int abcb200_set_tie_order(abcb200_ctx* ctx, int mode);
/* The host step of mode 1 on its own (no GPU involved): dist (N distances), order (top_n indices in ascending distance, ties in any
 * order) -> order rewritten as the first top_n entries of PLS::ordered(dist) when exact ties reach the output. Returns 1 when the
 * order was re-derived, 0 when it was already unambiguous, < 0 on a bad argument. */
int abcb200_tie_order_stdsort(const double* dist, int64_t N, int64_t top_n, uint64_t* order);
/* Number of kernels this context has launched so far (bench.py's gpu_launches claim). */
uint64_t abcb200_launch_count(abcb200_ctx* ctx);
/* Number of signed-rank tests (PLS::wilcoxon inside optimal_num_components) that had to be sorted exactly because
 * their rank-sum bracket straddled the threshold; all others were decided from the bracket alone (diagnostic). */
uint64_t abcb200_exact_test_count(abcb200_ctx* ctx);
/* Counters: 0 kernels launched so far; 1 signed-rank tests of the last component selection (sum over responses of the
 * PRESS argmin index); 2 of those, tests that needed the 4096-bin bracket; 3 tests sorted exactly so far; 4 component loop
 * of the last PLS fit (1 = pls_defl_kernel, deflated Gram matrix on chip; 2 = pls_gram_kernel, operands streamed from L2
 * by one CTA; 3 = pls_wide.cu, wide predictor sets: three whole-GPU launches per component); 5 components per block of the
 * last ranking's pipelined fit + hold-out validation (0: the stages ran one after the other); 6 whether the context owns an SM
 * partition (green contexts: 8 SMs for the one-CTA component loop, the rest for the kernels that run beside it); 7 component
 * selections so far whose exact tests went through the radix sort instead of the fine-bin ranking (ABCB200_EXACT_RADIX set, or
 * the fine-bin level gave up: more than 128 ambiguous tests after level 2, or a fine bin of more than 16384 elements); 8 rankings
 * so far whose order was re-derived by std::sort on the host because exact distance ties reached the output (abcb200_set_tie_order 1). */
uint64_t abcb200_stat(abcb200_ctx* ctx, int which);
/* Pinned host memory for callers that want full-rate H2D/D2H through the host entry points. */
int abcb200_host_alloc(size_t bytes, void** out);
int abcb200_host_free(void* p);
/* Device time in ms of the last call of a stage, measured with CUDA events on the context's stream.
 * stage: 0 moments+zscore, 1 pls fit, 2 hold-out residuals+PRESS, 3 Wilcoxon selection, 4 projection+distance,
 * 5 ordering, 6 doubled variance, 7 weight update, 8 H2D, 9 D2H. Returns < 0 for an unknown stage. */
double abcb200_stage_ms(abcb200_ctx* ctx, int stage);
#define ABCB200_NSTAGES 10
/* Device time in ms of the last launch of one hot kernel (CUDA events on the context's stream), for roofline reports.
 * kernel: 0 pls_gram_kernel (PLS component loop), 1 Gram products X^T Y + X^T X (atb kernels), 2 screen1_kernel (Wilcoxon
 * level 1), 3 screen2_kernel (level 2), 4 press_chk_kernel, 5 xb_kernel<0> (hold-out scores), 6 xb_kernel<1> (projection +
 * distance), 7 weights main kernel (weights_dmma_kernel or weights_diff_kernel), 8 zscore_kernel (metrics),
 * 9 pls_defl_kernel<.., LOO> (batched leave-one-out refits). */
double abcb200_kernel_ms(abcb200_ctx* ctx, int kernel);
#define ABCB200_NKERNELS 10

/* ---- ABC::particle_ranking_PLS, src/AbcUtil.cpp:423-458 -------------------------------------
 * met: N x K metrics (PLS predictors), par: N x P parameters (PLS responses), target: K observed metrics.
 * training rows are the first round(N*training_fraction) rows. method: ABCB200_KERNEL_TYPE1 is the
 * reference's default. top_n: number of leading entries of the order wanted (0 or >= N: full order);
 * order_out must hold that many. Ties in distance are ordered by ascending particle index (default) or as the reference's
 * std::sort leaves them (abcb200_set_tie_order).
 * dist_out (N, nullable), n_comp_used_out (nullable), n_comp_per_y_out (P entries, nullable). */
int abcb200_rank_pls(abcb200_ctx* ctx, const double* met, int64_t ld_met, const double* par, int64_t ld_par,
                     int64_t N, int K, int P, const double* target, double training_fraction, int method,
                     int64_t top_n, uint64_t* order_out, double* dist_out, int* n_comp_used_out,
                     int32_t* n_comp_per_y_out);
int abcb200_rank_pls_dev(abcb200_ctx* ctx, const double* met, int64_t ld_met, const double* par, int64_t ld_par,
                         int64_t N, int K, int P, const double* target, double training_fraction, int method,
                         int64_t top_n, uint64_t* order_out, double* dist_out, int* n_comp_used_out /*host*/,
                         int32_t* n_comp_per_y_out /*host*/);

/* ---- ABC::particle_ranking_simple, src/AbcUtil.cpp:408-421 ---------------------------------- */
int abcb200_rank_simple(abcb200_ctx* ctx, const double* met, int64_t ld_met, int64_t N, int K,
                        const double* target, int64_t top_n, uint64_t* order_out, double* dist_out);
int abcb200_rank_simple_dev(abcb200_ctx* ctx, const double* met, int64_t ld_met, int64_t N, int K,
                            const double* target, int64_t top_n, uint64_t* order_out, double* dist_out);

/* ---- ABC::calculate_doubled_variance, src/AbcUtil.cpp:528-537 (+ RunningStat.h:16-46) --------
 * params: n x P (the predictive prior's rows, gathered in rank order); dv_out: P. */
int abcb200_doubled_variance(abcb200_ctx* ctx, const double* params, int64_t ld, int64_t n, int P, double* dv_out);
int abcb200_doubled_variance_dev(abcb200_ctx* ctx, const double* params, int64_t ld, int64_t n, int P, double* dv_out);
/* Same, over rows params[idx[0..n-1], :] gathered on the device (AbcSmc.cpp:1045 fancy indexing). */
int abcb200_doubled_variance_gather_dev(abcb200_ctx* ctx, const double* params, int64_t ld, const uint64_t* idx,
                                        int64_t n, int P, double* gathered_out /* n x P, ld n, nullable */, double* dv_out);

/* ---- ABC::sample_predictive_priors, src/AbcUtil.cpp:378-390 (next-set proposals) --------
 * = ABC::sample_posterior (:366-376: weighted draw of rows via ABC::gsl_rng_nonuniform_int :111-121, P(row j) = w_j / sum w)
 * + ABC::gsl_ran_trunc_normal (:146-158) = Prior::noise per parameter (include/AbcSmc/Priors.h:18-41):
 * recast(theta[j,p] + sqrt(dv[p]) * N(0,1)), redrawn until valid, at most max_attempts times (reference: 1000), else the
 * prior's mean. The caller flattens its Parameter objects: valid(v) <=> lo[p] <= v <= hi[p] (+-inf for a Gaussian prior),
 * integral[p] != 0 rounds (DiscreteUniformPrior::recast, Priors.h:80; NULL = none), prior_mean[p] = Prior::get_mean().
 * Parity with the reference is DISTRIBUTIONAL (it consumes a gsl_rng stream; here Philox-4x32-10 counters keyed by seed:
 * the result depends on seed only). out: num_samples x P column-major; parent_out (nullable): the row drawn for each
 * sample; fallbacks_out (nullable): how many (sample, parameter) pairs fell back to the prior mean.
 * Errors: negative / non-finite / all-zero weights, negative variance (EINVAL; gsl_ran_discrete_preproc would abort). */
int abcb200_sample_predictive_priors(abcb200_ctx* ctx, uint64_t seed, int64_t num_samples, const double* weights, const double* theta,
                                     int64_t ld, int64_t n_pp, int P, const double* dv, const double* lo, const double* hi,
                                     const int32_t* integral, const double* prior_mean, int max_attempts, double* out, int64_t ld_out,
                                     uint64_t* parent_out, uint64_t* fallbacks_out);
/* Device pointers throughout (fallbacks_out: one device uint64_t, nullable); weights are not validated. */
int abcb200_sample_predictive_priors_dev(abcb200_ctx* ctx, uint64_t seed, int64_t num_samples, const double* weights, const double* theta,
                                         int64_t ld, int64_t n_pp, int P, const double* dv, const double* lo, const double* hi,
                                         const int32_t* integral, const double* prior_mean, int max_attempts, double* out, int64_t ld_out,
                                         uint64_t* parent_out, uint64_t* fallbacks_out);

/* ---- multivariate noise: ABC::setup_mvn_sampler src/AbcUtil.cpp:462-488, ABC::sample_mvn_predictive_priors :392-404 with
 * ABC::gsl_ran_trunc_mv_normal :123-144 (NOISE::MULTIVARIATE, AbcSmc.cpp:491-503) -------------------------------------------
 * setup: L_out (P x P, column-major, host) = lower Cholesky factor of the sample covariance of theta's rows (divisor n - 1)
 * with its diagonal doubled; EINVAL when that matrix is not positive definite (gsl_linalg_cholesky_decomp1 -> GSL_EDOM).
 * sample: parent row drawn as above, proposal = recast(parent + L z), z ~ N(0, I); the whole vector is redrawn until every
 * parameter lies in [lo, hi] (checked in order, as the reference). The reference retries without limit; here at most
 * max_attempts times, after which the sample keeps the (recast) parent row and is counted in failures_out (nullable).
 * Distributional parity, as for the independent variant. P <= 128. */
int abcb200_setup_mvn_sampler(abcb200_ctx* ctx, const double* theta, int64_t ld, int64_t n_pp, int P, double* L_out);
int abcb200_sample_mvn_predictive_priors(abcb200_ctx* ctx, uint64_t seed, int64_t num_samples, const double* weights, const double* theta,
                                         int64_t ld, int64_t n_pp, int P, const double* L, const double* lo, const double* hi,
                                         const int32_t* integral, int max_attempts, double* out, int64_t ld_out, uint64_t* parent_out,
                                         uint64_t* failures_out);

/* ---- ABC::weight_predictive_prior, src/AbcUtil.cpp:539-545 (set 0) and :547-586 (set > 0) ----
 * numer[i] = prod_p prior_p.likelihood(theta_new[i,p]) is computed by the caller (virtual call on the
 * host, src/AbcUtil.cpp:559-561; NULL means all ones). w_out: N_new L2-normalised weights.
 * algo: 0 auto, 1 pairwise-difference kernel (the reference's formulation), 2 DMMA inner-product kernel. */
int abcb200_weights_set0(abcb200_ctx* ctx, int64_t n, double* w_out /* host */);
int abcb200_weights(abcb200_ctx* ctx, const double* numer, const double* theta_new, int64_t ld_new, int64_t N_new,
                    const double* theta_old, int64_t ld_old, int64_t N_old, const double* w_old,
                    const double* dv_old, int P, int algo, double* w_out);
int abcb200_weights_dev(abcb200_ctx* ctx, const double* numer, const double* theta_new, int64_t ld_new, int64_t N_new,
                        const double* theta_old, int64_t ld_old, int64_t N_old, const double* w_old,
                        const double* dv_old, int P, int algo, double* w_out);
/* Sharded form: un-normalised weights for a slice of new-particle rows plus the slice's
 * sum of squares (device scalar); the caller all-reduces the sums and calls abcb200_scale_weights_dev. */
int abcb200_weights_unnorm_dev(abcb200_ctx* ctx, const double* numer, const double* theta_new, int64_t ld_new,
                               int64_t n_rows, const double* theta_old, int64_t ld_old, int64_t N_old,
                               const double* w_old, const double* dv_old, int P, int algo, double* w_out,
                               double* sumsq_out);
/* w[i] /= sqrt(*sumsq) when *sumsq > 0 (Eigen normalize() semantics, src/AbcUtil.cpp:583). */
int abcb200_scale_weights_dev(abcb200_ctx* ctx, double* w, int64_t n, const double* sumsq);

/* ---- the weight update sharded over the GPUs of one box (SURVEY.md §8 row e; caller: src/AbcSmc.cpp:1053-1064) -------------
 * New-particle rows are split over the members of a group: member r owns rows [r per, (r+1) per), per = ceil(N_new / G). The
 * previous set is broadcast GPU to GPU (ncclBroadcast), the conditioning maximum that picks the kernel is max-reduced so that
 * every member takes the same formulation for any G, the squared norm is sum-reduced (one double each), the slices are
 * all-gathered (device flavour) or copied to their place in the host result (host flavour). NCCL is loaded with dlopen at
 * first use; ABCB200_ENODEV when it cannot be.
 * A group is either ONE host process driving n GPUs (abcb200_group_create: what a C++ AbcSmc host uses; it owns one
 * context per GPU) or one member per process (abcb200_group_create_rank: launchers that start a process per GPU; the
 * 128-byte id comes from abcb200_group_unique_id on one rank and is passed round by the caller, e.g. over MPI). */
typedef struct abcb200_group abcb200_group;
int abcb200_group_create(int n_gpus /* <= 0: all visible */, const int* device_ids /* NULL: 0..n-1 */, abcb200_group** out);
int abcb200_group_unique_id(void* id_out, size_t bytes /* >= 128 */);
int abcb200_group_create_rank(abcb200_ctx* ctx, const void* id, int rank, int world, abcb200_group** out);
int abcb200_group_destroy(abcb200_group* g);
int abcb200_group_size(const abcb200_group* g);        /* members over all processes */
int abcb200_group_local_size(const abcb200_group* g);  /* members this process drives */
abcb200_ctx* abcb200_group_ctx(abcb200_group* g, int local_index);
const char* abcb200_group_last_error(abcb200_group* g);
/* ABC::weight_predictive_prior (set > 0) with HOST buffers, arguments as abcb200_weights; single-process groups. Each GPU
 * receives only its rows of the new set; the previous set crosses PCIe once. */
int abcb200_weights_sharded(abcb200_group* g, const double* numer, const double* theta_new, int64_t ld_new, int64_t N_new,
                            const double* theta_old, int64_t ld_old, int64_t N_old, const double* w_old,
                            const double* dv_old, int P, int algo, double* w_out);
/* Same with DEVICE buffers, called by every process of a process-per-GPU group (asynchronous on the member's stream). Every
 * member passes the full theta_new / numer; theta_old, w_old, dv_old are overwritten by the broadcast from member
 * bcast_root when bcast_root >= 0 (pass -1 when every member already holds them). w_gathered (N_new, nullable): all
 * weights, on every member; w_slice_out (per = ceil(N_new / G) entries, nullable): this member's rows. */
int abcb200_weights_sharded_dev(abcb200_group* g, const double* numer, const double* theta_new, int64_t ld_new, int64_t N_new,
                                double* theta_old, int64_t ld_old, int64_t N_old, double* w_old, double* dv_old, int P,
                                int algo, int bcast_root, double* w_gathered, double* w_slice_out);

/* ---- one call per SMC set, the previous set resident on the device -------------------------------------
 * = the body of the set loop of AbcSmc::read_SMC_sets_from_database (src/AbcSmc.cpp:634-664): filter (0 = particle_ranking_PLS,
 * 1 = particle_ranking_simple), keep the first top_n, gather their rows, AbcLog::filtering_report's statistics (src/AbcLog.cpp:81-124),
 * then AbcSmc::calculate_predictive_prior_weights (:1041-1066): doubled variance and weights against set t-1. The chain keeps set
 * t-1's gathered parameters, weights and doubled variance in device memory between calls (the reference recomputes them for every
 * earlier set on each --process run, :664); abcb200_chain_state reads them back for the host to persist, abcb200_chain_restore
 * re-seeds a chain from persisted values instead of replaying the sets.
 * Numerator prod_p prior_p.likelihood(theta) (src/AbcUtil.cpp:559-561), in this order of preference: numer_all (N host values, one per
 * particle of the set, for hosts with their own Parameter classes), or flat priors evaluated on the device (prior_type[p]: 0
 * ContinuousUniformPrior [a, b], 1 DiscreteUniformPrior [a, b], 2 GaussianPrior mean a, sd b; include/AbcSmc/Priors.h:44-110), or
 * both NULL: 1. order_out: top_n indices; weights_out: top_n (set 0: 1 / top_n, un-normalised, :539-545); dv_out: P;
 * report_out (nullable): 1 + 2 (P + K) doubles = NRMSE of the posterior metric means (ABC::calculate_nrmse), posterior means of
 * the P parameters and K metrics, posterior medians of the same (ABC::median). */
typedef struct abcb200_chain abcb200_chain;
int abcb200_chain_create(abcb200_ctx* ctx, int P, abcb200_chain** out);
int abcb200_chain_destroy(abcb200_chain* ch);
int abcb200_chain_sets(const abcb200_chain* ch);
int abcb200_chain_nparams(const abcb200_chain* ch);   /* the P the chain was created with (0 for a null chain) */
int abcb200_chain_process_set(abcb200_chain* ch, const double* met, int64_t ld_met, const double* par, int64_t ld_par, int64_t N, int K,
                              const double* target, int filter, double training_fraction, int method, int64_t top_n,
                              const int32_t* prior_type, const double* prior_a, const double* prior_b, const double* numer_all,
                              uint64_t* order_out, double* weights_out, double* dv_out, double* report_out, int* n_comp_used_out);
int abcb200_chain_state(abcb200_chain* ch, int64_t* n_out, double* theta_out, int64_t ld_out, double* weights_out, double* dv_out);
int abcb200_chain_restore(abcb200_chain* ch, const double* theta, int64_t ld, int64_t n, const double* weights, const double* dv, int sets_done);

/* ---- storage boundary: AbcSmc's SQLite job database, host code only -----------------------------------------
 * Schema src/AbcSmc.cpp:819-834 (tables job, par, met). Replaces the per-field copy of the three-table join through sqdb
 * (:596-621) by three scans merged by serial and written straight into column-major host buffers (row = particleIdx), and the one-UPDATE-string-per-
 * particle rank write-back (:653-661) by one prepared UPDATE in one transaction. libsqlite3.so.0 is loaded with dlopen at first use
 * (ABCB200_ENODEV when absent). Errors: abcb200_db_last_error() (thread-local message). */
const char* abcb200_db_last_error(void);
int abcb200_db_set_shape(const char* db_path, int set, int64_t* n_out, int* npar_out, int* nmet_out);
/* par: N x P (ld_par), met: N x K (ld_met); serial_out (N, nullable): job.serial per particle; posterior_out (N, nullable): job.posterior
 * (-1: not ranked). EINVAL when particleIdx is not 0 .. N-1 or a metric is NULL. Use abcb200_host_alloc buffers for full-rate H2D. */
int abcb200_db_load_set(const char* db_path, int set, int64_t N, int P, int K, double* par, int64_t ld_par, double* met, int64_t ld_met,
                        int64_t* serial_out, int32_t* posterior_out);
/* job.posterior = i for the particle whose serial is serial_by_rank[i], i < n */
int abcb200_db_write_ranks(const char* db_path, const int64_t* serial_by_rank, int64_t n);
/* `--process` of one not-yet-ranked set in one call: bulk load into pinned buffers -> abcb200_chain_process_set -> rank write-back.
 * Outputs as abcb200_chain_process_set (order_out nullable here). */
int abcb200_chain_process_db_set(abcb200_chain* ch, const char* db_path, int set, const double* target, int filter, double training_fraction,
                                 int method, int64_t top_n, const int32_t* prior_type, const double* prior_a, const double* prior_b,
                                 uint64_t* order_out, double* weights_out, double* dv_out, double* report_out, int* n_comp_used_out);

/* ---- free functions of namespace PLS / ABC --------------------------------------------------- */
/* PLS::colwise_stdev + colwise mean, lib/PLS/src/pls.cpp:69-87 */
int abcb200_colwise_moments(abcb200_ctx* ctx, const double* X, int64_t ld, int64_t N, int K, double* mean_out, double* sd_out);
/* PLS::colwise_z_scores, lib/PLS/src/pls.cpp:93-111 (mean/sd NULL: computed from X) */
int abcb200_colwise_z_scores(abcb200_ctx* ctx, const double* X, int64_t ld, int64_t N, int K, const double* mean,
                             const double* sd, double* Z_out, int64_t ld_out);
/* The two products PLS::Model::plsr starts from, lib/PLS/src/pls.cpp:396 (XY = X^T Y) and :398 (XX = X^T X, KERNEL_TYPE2):
   XX_out K x K (ld K, both triangles), XY_out K x M (ld K). One pass over X, FP64 tensor cores. */
int abcb200_gram(abcb200_ctx* ctx, const double* X, int64_t ldx, const double* Y, int64_t ldy, int64_t N, int K, int M,
                 double* XX_out, double* XY_out);
/* ABC::euclidean, src/AbcUtil.cpp:320-324 */
int abcb200_euclidean(abcb200_ctx* ctx, const double* S, int64_t ld, int64_t N, int K, const double* ref, double* out);
/* PLS::ordered, lib/PLS/include/PLS/pls.h:58-69 (ties: ascending index, or std::sort's placement: abcb200_set_tie_order) */
int abcb200_ordered(abcb200_ctx* ctx, const double* v, int64_t n, uint64_t* order_out);
/* The first top_n entries of PLS::ordered(v) only (what AbcSmc.cpp:645-646 keeps): radix select + small sort. */
int abcb200_ordered_top(abcb200_ctx* ctx, const double* v, int64_t n, int64_t top_n, uint64_t* order_out);
/* PLS::wilcoxon, lib/PLS/src/pls.cpp:190-211 */
int abcb200_wilcoxon(abcb200_ctx* ctx, const double* err1, const double* err2, int64_t n, double* p_out);

/* ---- PLS::Model, lib/PLS/include/PLS/pls.h:184-266, lib/PLS/src/pls.cpp:340-510 --------------- */
/* Model(X, Y, algorithm, max_components): fits immediately (pls.cpp:340-353). X: N x K, Y: N x M. */
int abcb200_pls_fit(abcb200_ctx* ctx, const double* X, int64_t ldx, const double* Y, int64_t ldy, int64_t N, int K,
                    int M, int method, int max_components, abcb200_pls** out);
int abcb200_pls_free(abcb200_pls* m);
/* which: 'P','W','R' (K x A), 'Q' (M x A), 'T' (N x A, KERNEL_TYPE1 only). Real parts; out is column-major, tight. */
int abcb200_pls_get(abcb200_pls* m, char which, double* out);
/* Model::scores, pls.cpp:439-442. out: n x comp */
int abcb200_pls_scores(abcb200_pls* m, const double* Xnew, int64_t ld, int64_t n, int comp, double* out);
/* Model::coefficients, pls.cpp:444-447. out: K x M */
int abcb200_pls_coefficients(abcb200_pls* m, int comp, double* out);
/* Model::fitted_values / residuals / SSE, pls.cpp:449-459. out: n x M / n x M / M */
int abcb200_pls_fitted_values(abcb200_pls* m, const double* Xnew, int64_t ld, int64_t n, int comp, double* out);
int abcb200_pls_residuals(abcb200_pls* m, const double* Xnew, int64_t ldx, const double* Ynew, int64_t ldy, int64_t n, int comp, double* out);
int abcb200_pls_sse(abcb200_pls* m, const double* Xnew, int64_t ldx, const double* Ynew, int64_t ldy, int64_t n, int comp, double* out);
/* Model::explained_variance, pls.cpp:461-467: 1 - SSE / SST(Y_new). out: M */
int abcb200_pls_explained_variance(abcb200_pls* m, const double* Xnew, int64_t ldx, const double* Ynew, int64_t ldy, int64_t n, int comp, double* out);
/* Model::cv_NEW_DATA + PLS::validation + PLS::optimal_num_components (pls.cpp:494-510, 235-289), streamed:
 * the M x n x A error cube is never materialised. press_out: M x A (column-major, nullable), out_type RESS|MSE;
 * n_comp_out: M component counts (nullable). */
int abcb200_pls_cv_new_data(abcb200_pls* m, const double* Xnew, int64_t ldx, const double* Ynew, int64_t ldy, int64_t n,
                            int out_type, double alpha, double* press_out, int32_t* n_comp_out);

/* PLS::validation + PLS::optimal_num_components (pls.cpp:235-261, 265-289) on a materialised PLS::Residual
 * (pls.h:44-53): errors[(y * A + c) * n + i] = error of response y, row i, predicted with c + 1 components (one n x A
 * column-major matrix per response, as Residual::errors() holds them). press_out: M x A column-major (nullable);
 * n_comp_out: M counts (nullable). */
int abcb200_residual_select(abcb200_ctx* ctx, const double* errors, int64_t n, int M, int A, int out_type, double alpha,
                            double* press_out, int32_t* n_comp_out);
/* Model::cv_LOO, pls.cpp:469-491 (+ validation / optimal_num_components on its result). X (N x K), Y (N x M) are the
 * matrices the model holds (_X, _Y); A = the model's component count (every refit uses min(K, A) components, :477, :479).
 * The N refits run as rank-one down-dates of X^T X, X^T Y, batched over the SMs. errors_out: M x (N x A) cube in the
 * layout above (nullable). */
int abcb200_pls_cv_loo(abcb200_ctx* ctx, const double* X, int64_t ldx, const double* Y, int64_t ldy, int64_t N, int K, int M,
                       int A, int out_type, double alpha, double* errors_out, double* press_out, int32_t* n_comp_out);
/* Model::cv_LSO, pls.cpp:512-549. The random splits stay with the caller (std::shuffle on a std::mt19937 is
 * libstdc++-specific, rand_nchoosek :218-227): shuffles holds, per trial, the N shuffled row indices `full`; the first
 * N - test_size train, the rest are predicted, in that order. errors_out: M x (num_trials * test_size x A) (nullable). */
int abcb200_pls_cv_lso(abcb200_ctx* ctx, const double* X, int64_t ldx, const double* Y, int64_t ldy, int64_t N, int K, int M,
                       int A, int method, const uint64_t* shuffles, int64_t test_size, int64_t num_trials, int out_type,
                       double alpha, double* errors_out, double* press_out, int32_t* n_comp_out);

#ifdef __cplusplus
}
#endif
#endif /* ABCSMC_B200_H */
