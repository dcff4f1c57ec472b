// abc_b200.hpp — header-only C++17 host adapter: the reference's own signatures for the per-set hot path, implemented
// on the C ABI of libabcsmc_b200.so (include/abcsmc_b200.h). No numerics happen here: every function stages its
// arguments and calls the CUDA library; there is no CPU fallback (a missing GPU is a fatal error, as every error on
// this path is in the reference: message on std::cerr, then exit()).
//
// Mirrors (names, argument order and meaning identical):
//   ABC::particle_ranking_PLS / particle_ranking_simple   include/AbcSmc/AbcUtil.h:146-155, src/AbcUtil.cpp:408-458
//   ABC::weight_predictive_prior (both overloads)          include/AbcSmc/AbcUtil.h:157-168, src/AbcUtil.cpp:539-586
//   ABC::calculate_doubled_variance                        include/AbcSmc/AbcUtil.h:170-172, src/AbcUtil.cpp:528-537
//   ABC::sample_predictive_priors (next-set proposals)     include/AbcSmc/AbcUtil.h:121-126, src/AbcUtil.cpp:378-390
//   ABC::euclidean                                         include/AbcSmc/AbcUtil.h:103,     src/AbcUtil.cpp:320-324
//   PLS::ordered, colwise_stdev, colwise_z_scores, wilcoxon, optimal_num_components (on a streamed validation),
//   PLS::Model                                             lib/PLS/include/PLS/pls.h:58-69, 97-159, 184-266
//
// The matrix / vector types are template parameters so that the header compiles with or without Eigen. What it needs
// from them is what Eigen::MatrixXd / RowVectorXd / VectorXd provide: column-major storage behind data(), rows(),
// cols(), outerStride() for matrices; data(), size() and a size constructor for vectors; a (rows, cols) constructor
// for matrices. `Parameter` only needs `double likelihood(double) const` (include/AbcSmc/Parameter.h:58).
//
// Inside AbcSmc, define ABCB200_DROP_IN before including this header *instead of* compiling the five function bodies in
// src/AbcUtil.cpp: it then defines them in namespace ABC with the reference's exact types.
#ifndef ABC_B200_HPP
#define ABC_B200_HPP

#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <cmath>
#include <limits>
#include <algorithm>
#include <vector>

#include "../../include/abcsmc_b200.h"

namespace abcb200 {

// One library context per host thread (the reference is single threaded; AbcSmc.cpp:452-559 runs in one process).
class Context {
  public:
    static Context& instance(int device = -1) {
        static thread_local Context ctx(device);
        return ctx;
    }
    abcb200_ctx* handle() const { return h_; }
    // Every failure on this path is fatal in the reference (assert or cerr + exit); keep that contract.
    void check(int rc, const char* where) const {
        if (rc == ABCB200_OK) return;
        std::cerr << "ERROR: " << where << " failed (" << rc << "): " << abcb200_last_error(h_) << std::endl;
        std::exit(-300 + rc);
    }

  private:
    explicit Context(int device) : h_(nullptr) {
        if (device < 0) { const char* e = std::getenv("ABCB200_DEVICE"); device = e ? std::atoi(e) : 0; }
        const int rc = abcb200_create(device, &h_);
        if (rc != ABCB200_OK) {
            std::cerr << "ERROR: abcb200_create(" << device << ") failed (" << rc << "): no usable sm_100 CUDA device; this build has no CPU path" << std::endl;
            std::exit(-300 + rc);
        }
    }
    ~Context() { abcb200_destroy(h_); }
    Context(const Context&) = delete;
    abcb200_ctx* h_;
};

template <class M> inline int64_t ld(const M& m) { return (int64_t)m.outerStride(); }

}  // namespace abcb200

namespace ABC_B200 {

// std::vector<size_t> ABC::particle_ranking_PLS(X_orig = metrics, Y_orig = parameters, target_values, training_fraction)
// Returns the full order, as the reference does (AbcUtil.cpp:457); AbcSmc then keeps the first next_pred_prior_size
// entries (AbcSmc.cpp:645-646) — pass top_n to receive only those.
template <class Mat2D, class Row>
std::vector<size_t> particle_ranking_PLS(const Mat2D& X_orig, const Mat2D& Y_orig, const Row& target_values, const double training_fraction,
                                         const size_t top_n = 0) {
    auto& c = abcb200::Context::instance();
    const int64_t N = (int64_t)X_orig.rows();
    const size_t n_out = (top_n == 0 || top_n > (size_t)N) ? (size_t)N : top_n;
    std::vector<uint64_t> order(n_out);
    int used = 0;
    c.check(abcb200_rank_pls(c.handle(), X_orig.data(), abcb200::ld(X_orig), Y_orig.data(), abcb200::ld(Y_orig), N, (int)X_orig.cols(), (int)Y_orig.cols(),
                             target_values.data(), training_fraction, ABCB200_KERNEL_TYPE1, (int64_t)n_out, order.data(), nullptr, &used, nullptr),
            "particle_ranking_PLS");
    return std::vector<size_t>(order.begin(), order.end());
}

// std::vector<size_t> ABC::particle_ranking_simple(X_orig, Y_orig (unused, as in the reference), target_values)
template <class Mat2D, class Row>
std::vector<size_t> particle_ranking_simple(const Mat2D& X_orig, const Mat2D& /*Y_orig*/, const Row& target_values, const size_t top_n = 0) {
    auto& c = abcb200::Context::instance();
    const int64_t N = (int64_t)X_orig.rows();
    const size_t n_out = (top_n == 0 || top_n > (size_t)N) ? (size_t)N : top_n;
    std::vector<uint64_t> order(n_out);
    c.check(abcb200_rank_simple(c.handle(), X_orig.data(), abcb200::ld(X_orig), N, (int)X_orig.cols(), target_values.data(), (int64_t)n_out, order.data(), nullptr),
            "particle_ranking_simple");
    return std::vector<size_t>(order.begin(), order.end());
}

// Row ABC::calculate_doubled_variance(params)
template <class Row, class Mat2D>
Row calculate_doubled_variance(const Mat2D& params) {
    auto& c = abcb200::Context::instance();
    Row dv(params.cols());
    c.check(abcb200_doubled_variance(c.handle(), params.data(), abcb200::ld(params), (int64_t)params.rows(), (int)params.cols(), dv.data()), "calculate_doubled_variance");
    return dv;
}

// Row ABC::weight_predictive_prior(mpars, params): set 0, uniform 1 / N (AbcUtil.cpp:539-545)
template <class Row, class Parameter, class Mat2D>
Row weight_predictive_prior(const std::vector<const Parameter*>& /*mpars*/, const Mat2D& params) {
    auto& c = abcb200::Context::instance();
    Row w(params.rows());
    c.check(abcb200_weights_set0(c.handle(), (int64_t)params.rows(), w.data()), "weight_predictive_prior (set 0)");
    return w;
}

// Row ABC::weight_predictive_prior(mpars, params, prev_params, prev_weights, prev_doubled_variance) (AbcUtil.cpp:547-586)
// The numerator prod_p mpars[p]->likelihood(params(i, p)) stays on the host behind the virtual call (:559-561).
template <class Row, class Parameter, class Mat2D, class RowIn>
Row weight_predictive_prior(const std::vector<const Parameter*>& mpars, const Mat2D& params, const Mat2D& prev_params, const RowIn& prev_weights,
                            const RowIn& prev_doubled_variance) {
    auto& c = abcb200::Context::instance();
    const int64_t n = (int64_t)params.rows(), ldp = abcb200::ld(params);
    const int P = (int)params.cols();
    std::vector<double> numer((size_t)n, 1.0);
    for (int p = 0; p < P; p++) {
        const double* col = params.data() + (int64_t)p * ldp;
        for (int64_t i = 0; i < n; i++) numer[(size_t)i] *= mpars[(size_t)p]->likelihood(col[i]);
    }
    Row w(n);
    c.check(abcb200_weights(c.handle(), numer.data(), params.data(), ldp, n, prev_params.data(), abcb200::ld(prev_params), (int64_t)prev_params.rows(),
                            prev_weights.data(), prev_doubled_variance.data(), P, 0, w.data()),
            "weight_predictive_prior");
    return w;
}

// What Prior::noise needs of a Parameter, flattened for the device (include/AbcSmc/Priors.h:18-41): the validity interval
// [lo, hi] (valid(v) <=> likelihood(v) != 0, Parameter.h:77), whether recast() rounds (DiscreteUniformPrior, Priors.h:80)
// and the prior mean. The reference's Parameter has no accessor for its bounds, so they are recovered from get_mean() /
// get_sd() (uniform priors: half-width = sd * sqrt(3), Priors.h:66, 92) and then walked to the exact boundary doubles with
// valid(); a prior that is still valid 1.8 sd from its mean on both sides is unbounded (GaussianPrior).
struct FlatPrior { double lo, hi, mean; int32_t integral; };
template <class Parameter>
FlatPrior flatten_prior(const Parameter& par) {
    FlatPrior f;
    f.mean = (double)par.get_mean();
    const double sd = (double)par.get_sd();
    f.integral = (par.recast(0.5) != 0.5) ? 1 : 0;
    const double half = sd * std::sqrt(3.0);
    f.lo = f.mean - half; f.hi = f.mean + half;
    // the only rounding prior of the reference is the discrete uniform one (Priors.h:60-84): bounded, integer bounds. (It must
    // not go through the probe below: recast() pulls points up to 0.5 outside the range back onto its end points.)
    if (f.integral) { f.lo = std::round(f.lo); f.hi = std::round(f.hi); return f; }
    if (par.valid(f.mean + 1.8 * sd) && par.valid(f.mean - 1.8 * sd)) {
        f.lo = -std::numeric_limits<double>::infinity(); f.hi = std::numeric_limits<double>::infinity();
        return f;
    }
    const double inf = std::numeric_limits<double>::infinity();
    for (int i = 0; i < 64 && !par.valid(f.lo); i++) f.lo = std::nextafter(f.lo, inf);
    for (int i = 0; i < 64 && par.valid(std::nextafter(f.lo, -inf)); i++) f.lo = std::nextafter(f.lo, -inf);
    for (int i = 0; i < 64 && !par.valid(f.hi); i++) f.hi = std::nextafter(f.hi, -inf);
    for (int i = 0; i < 64 && par.valid(std::nextafter(f.hi, inf)); i++) f.hi = std::nextafter(f.hi, inf);
    return f;
}

// Mat2D ABC::sample_predictive_priors(RNG, num_samples, weights, parameter_prior, pars, doubled_variance) (AbcUtil.cpp:378-390),
// the gsl_rng replaced by a 64-bit seed (draw it from the caller's generator; include/abcsmc_b200.h: distributional parity).
// Like the reference it reports prior-mean fall-backs on stderr (Priors.h:27).
template <class Mat2D, class Parameter, class Col, class Row>
Mat2D sample_predictive_priors(uint64_t seed, const size_t num_samples, const Col& weights, const Mat2D& parameter_prior,
                               const std::vector<const Parameter*>& pars, const Row& doubled_variance) {
    auto& c = abcb200::Context::instance();
    const int P = (int)parameter_prior.cols();
    std::vector<double> lo((size_t)P), hi((size_t)P), mean((size_t)P);
    std::vector<int32_t> integral((size_t)P);
    for (int p = 0; p < P; p++) {
        const FlatPrior f = flatten_prior(*pars[(size_t)p]);
        lo[(size_t)p] = f.lo; hi[(size_t)p] = f.hi; mean[(size_t)p] = f.mean; integral[(size_t)p] = f.integral;
    }
    Mat2D out((long)num_samples, (long)P);
    uint64_t fallbacks = 0;
    c.check(abcb200_sample_predictive_priors(c.handle(), seed, (int64_t)num_samples, weights.data(), parameter_prior.data(), abcb200::ld(parameter_prior),
                                             (int64_t)parameter_prior.rows(), P, doubled_variance.data(), lo.data(), hi.data(), integral.data(), mean.data(),
                                             1000, out.data(), abcb200::ld(out), nullptr, &fallbacks),
            "sample_predictive_priors");
    if (fallbacks) std::cerr << "ERROR: failed to draw valid noise from prior for " << fallbacks << " value(s) - returning mean value." << std::endl;
    return out;
}

// Col ABC::euclidean(sims, ref)
template <class Col, class Mat2D, class Row>
Col euclidean(const Mat2D& sims, const Row& ref) {
    auto& c = abcb200::Context::instance();
    Col d(sims.rows());
    c.check(abcb200_euclidean(c.handle(), sims.data(), abcb200::ld(sims), (int64_t)sims.rows(), (int)sims.cols(), ref.data(), d.data()), "euclidean");
    return d;
}

}  // namespace ABC_B200

namespace PLS_B200 {

typedef enum { KERNEL_TYPE1 = ABCB200_KERNEL_TYPE1, KERNEL_TYPE2 = ABCB200_KERNEL_TYPE2 } METHOD;
typedef enum { RESS = ABCB200_RESS, MSE = ABCB200_MSE } VALIDATION_OUTPUT;

// std::vector<size_t> PLS::ordered(v) — ties in ascending index (the reference's tie order is unspecified)
template <class Vec>
std::vector<size_t> ordered(const Vec& v) {
    auto& c = abcb200::Context::instance();
    std::vector<uint64_t> o((size_t)v.size());
    c.check(abcb200_ordered(c.handle(), v.data(), (int64_t)v.size(), o.data()), "ordered");
    return std::vector<size_t>(o.begin(), o.end());
}
template <class Row, class Mat2D>
Row colwise_stdev(const Mat2D& mat) {
    auto& c = abcb200::Context::instance();
    Row sd(mat.cols());
    c.check(abcb200_colwise_moments(c.handle(), mat.data(), abcb200::ld(mat), (int64_t)mat.rows(), (int)mat.cols(), nullptr, sd.data()), "colwise_stdev");
    return sd;
}
template <class Mat2D>
Mat2D colwise_z_scores(const Mat2D& mat) {
    auto& c = abcb200::Context::instance();
    Mat2D z(mat.rows(), mat.cols());
    c.check(abcb200_colwise_z_scores(c.handle(), mat.data(), abcb200::ld(mat), (int64_t)mat.rows(), (int)mat.cols(), nullptr, nullptr, z.data(), abcb200::ld(z)), "colwise_z_scores");
    return z;
}
template <class Mat2D, class Row>
Mat2D colwise_z_scores(const Mat2D& mat, const Row& mean, const Row& stdev) {
    auto& c = abcb200::Context::instance();
    Mat2D z(mat.rows(), mat.cols());
    c.check(abcb200_colwise_z_scores(c.handle(), mat.data(), abcb200::ld(mat), (int64_t)mat.rows(), (int)mat.cols(), mean.data(), stdev.data(), z.data(), abcb200::ld(z)),
            "colwise_z_scores");
    return z;
}
template <class Col>
double wilcoxon(const Col& err_1, const Col& err_2) {
    auto& c = abcb200::Context::instance();
    double p = 0;
    c.check(abcb200_wilcoxon(c.handle(), err_1.data(), err_2.data(), (int64_t)err_1.size(), &p), "wilcoxon");
    return p;
}

// What cv_NEW_DATA + validation + optimal_num_components produce, streamed on the device (the M x n x A error cube of
// PLS::Residual is never materialised): validation(RESS | MSE) as an M x A matrix and the component counts.
template <class Mat2D>
struct Validation {
    Mat2D press;                      // PLS::validation(residual, out_type), M x A
    std::vector<size_t> num_components;   // PLS::optimal_num_components(residual, ALPHA)
};

// struct PLS::Model (pls.h:184-266). Real parts only: every consumer of the reference's complex factors takes .real().
template <class Mat2D, class Row>
struct Model {
    Model(const Mat2D& X, const Mat2D& Y, const METHOD& algorithm, const size_t& max_components)
        : K_((int)X.cols()), M_((int)Y.cols()), A_((int)max_components), N_((int64_t)X.rows()), h_(nullptr) {
        auto& c = abcb200::Context::instance();
        c.check(abcb200_pls_fit(c.handle(), X.data(), abcb200::ld(X), Y.data(), abcb200::ld(Y), N_, K_, M_, (int)algorithm, A_, &h_), "PLS::Model");
    }
    Model(const Mat2D& X, const Mat2D& Y, const METHOD& algorithm = KERNEL_TYPE1) : Model(X, Y, algorithm, (size_t)X.cols()) {}
    ~Model() { abcb200_pls_free(h_); }
    Model(const Model&) = delete;
    Model& operator=(const Model&) = delete;

    const Mat2D scores(const Mat2D& X_new, const size_t comp) const {
        Mat2D out(X_new.rows(), comp);
        ck(abcb200_pls_scores(h_, X_new.data(), abcb200::ld(X_new), (int64_t)X_new.rows(), (int)comp, out.data()), "Model::scores");
        return out;
    }
    const Mat2D scores(const Mat2D& X_new) const { return scores(X_new, (size_t)A_); }
    const Mat2D loadingsX(const size_t comp) const { return factor('P', K_, comp); }   // declared, never defined, in the reference (pls.h:207-211)
    const Mat2D loadingsY(const size_t comp) const { return factor('Q', M_, comp); }
    const Mat2D coefficients(const size_t comp) const {
        Mat2D out(K_, M_);
        ck(abcb200_pls_coefficients(h_, (int)comp, out.data()), "Model::coefficients");
        return out;
    }
    const Mat2D coefficients() const { return coefficients((size_t)A_); }
    const Mat2D fitted_values(const Mat2D& X, const size_t comp) const {
        Mat2D out(X.rows(), M_);
        ck(abcb200_pls_fitted_values(h_, X.data(), abcb200::ld(X), (int64_t)X.rows(), (int)comp, out.data()), "Model::fitted_values");
        return out;
    }
    const Mat2D fitted_values(const Mat2D& X) const { return fitted_values(X, (size_t)A_); }
    const Mat2D residuals(const Mat2D& X, const Mat2D& Y, const size_t comp) const {
        Mat2D out(X.rows(), M_);
        ck(abcb200_pls_residuals(h_, X.data(), abcb200::ld(X), Y.data(), abcb200::ld(Y), (int64_t)X.rows(), (int)comp, out.data()), "Model::residuals");
        return out;
    }
    const Mat2D residuals(const Mat2D& X, const Mat2D& Y) const { return residuals(X, Y, (size_t)A_); }
    const Row SSE(const Mat2D& X, const Mat2D& Y, const size_t comp) const {
        Row out(M_);
        ck(abcb200_pls_sse(h_, X.data(), abcb200::ld(X), Y.data(), abcb200::ld(Y), (int64_t)X.rows(), (int)comp, out.data()), "Model::SSE");
        return out;
    }
    const Row SSE(const Mat2D& X, const Mat2D& Y) const { return SSE(X, Y, (size_t)A_); }
    // cv_NEW_DATA(X, Y) followed by validation(out_type) and optimal_num_components(ALPHA) (pls.cpp:494-510, 235-289)
    Validation<Mat2D> cv_NEW_DATA(const Mat2D& X, const Mat2D& Y, const VALIDATION_OUTPUT out_type = RESS, const double ALPHA = 0.1) const {
        Validation<Mat2D> v{Mat2D(M_, A_), std::vector<size_t>((size_t)M_)};
        std::vector<int32_t> nc((size_t)M_);
        ck(abcb200_pls_cv_new_data(h_, X.data(), abcb200::ld(X), Y.data(), abcb200::ld(Y), (int64_t)X.rows(), (int)out_type, ALPHA, v.press.data(), nc.data()), "Model::cv_NEW_DATA");
        for (int y = 0; y < M_; y++) v.num_components[(size_t)y] = (size_t)nc[(size_t)y];
        return v;
    }
    // cv_LOO (pls.cpp:469-491) followed by validation / optimal_num_components. The reference's Model keeps copies of X and Y
    // (_X, _Y, pls.cpp:344); this wrapper does not, so the matrices the model was built from are passed again.
    Validation<Mat2D> cv_LOO(const Mat2D& X, const Mat2D& Y, const VALIDATION_OUTPUT out_type = RESS, const double ALPHA = 0.1) const {
        Validation<Mat2D> v{Mat2D(M_, A_), std::vector<size_t>((size_t)M_)};
        std::vector<int32_t> nc((size_t)M_);
        ck(abcb200_pls_cv_loo(abcb200::Context::instance().handle(), X.data(), abcb200::ld(X), Y.data(), abcb200::ld(Y), (int64_t)X.rows(), K_, M_, A_,
                              (int)out_type, ALPHA, nullptr, v.press.data(), nc.data()), "Model::cv_LOO");
        for (int y = 0; y < M_; y++) v.num_components[(size_t)y] = (size_t)nc[(size_t)y];
        return v;
    }
    // cv_LSO (pls.cpp:512-549): the random splits are drawn here exactly as PLS::rand_nchoosek does (std::shuffle of the running
    // index vector on the caller's generator, pls.cpp:217-227), so a given generator state gives the reference's partitions.
    template <class RNG>
    Validation<Mat2D> cv_LSO(const Mat2D& X, const Mat2D& Y, const double test_fraction, const size_t num_trials, RNG& rng, const METHOD algorithm = KERNEL_TYPE1,
                             const VALIDATION_OUTPUT out_type = RESS, const double ALPHA = 0.1) const {
        const size_t N = (size_t)X.rows();
        const size_t test_size = (size_t)(test_fraction * (double)N + 0.5);
        std::vector<uint64_t> full(N), shuffles(N * num_trials);
        for (size_t i = 0; i < N; i++) full[i] = i;
        for (size_t t = 0; t < num_trials; t++) {
            std::shuffle(full.begin(), full.end(), rng);
            std::copy(full.begin(), full.end(), shuffles.begin() + (std::ptrdiff_t)(t * N));
        }
        Validation<Mat2D> v{Mat2D(M_, A_), std::vector<size_t>((size_t)M_)};
        std::vector<int32_t> nc((size_t)M_);
        ck(abcb200_pls_cv_lso(abcb200::Context::instance().handle(), X.data(), abcb200::ld(X), Y.data(), abcb200::ld(Y), (int64_t)N, K_, M_, A_, (int)algorithm,
                              shuffles.data(), (int64_t)test_size, (int64_t)num_trials, (int)out_type, ALPHA, nullptr, v.press.data(), nc.data()), "Model::cv_LSO");
        for (int y = 0; y < M_; y++) v.num_components[(size_t)y] = (size_t)nc[(size_t)y];
        return v;
    }

  private:
    void ck(int rc, const char* where) const { abcb200::Context::instance().check(rc, where); }
    Mat2D factor(char which, int rows, size_t comp) const {
        Mat2D full(rows, A_);
        ck(abcb200_pls_get(h_, which, full.data()), "Model factor");
        Mat2D out(rows, comp);
        for (size_t a = 0; a < comp; a++) for (int r = 0; r < rows; r++) out.data()[a * (size_t)out.outerStride() + r] = full.data()[a * (size_t)full.outerStride() + r];
        return out;
    }
    int K_, M_, A_;
    int64_t N_;
    abcb200_pls* h_;
};

}  // namespace PLS_B200

#ifdef ABCB200_DROP_IN
// Definitions (external linkage, include in exactly one .cpp) with the reference's exact names and types; compile
// this in place of the bodies in src/AbcUtil.cpp.
// Requires the reference's headers (Eigen typedefs Mat2D / Row / Col of lib/PLS/include/PLS/pls.h:22-27 and Parameter).
namespace ABC {
std::vector<size_t> particle_ranking_PLS(const Mat2D& X_orig, const Mat2D& Y_orig, const Row& target_values, const float_type training_fraction) {
    return ABC_B200::particle_ranking_PLS(X_orig, Y_orig, target_values, (double)training_fraction);
}
std::vector<size_t> particle_ranking_simple(const Mat2D& X_orig, const Mat2D& Y_orig, const Row& target_values) {
    return ABC_B200::particle_ranking_simple(X_orig, Y_orig, target_values);
}
Row calculate_doubled_variance(const Mat2D& params) { return ABC_B200::calculate_doubled_variance<Row>(params); }
Row weight_predictive_prior(const std::vector<const Parameter*>& mpars, const Mat2D& params) { return ABC_B200::weight_predictive_prior<Row>(mpars, params); }
Row weight_predictive_prior(const std::vector<const Parameter*>& mpars, const Mat2D& params, const Mat2D& prev_params, const Row& prev_weights,
                                   const Row& prev_doubled_variance) {
    return ABC_B200::weight_predictive_prior<Row>(mpars, params, prev_params, prev_weights, prev_doubled_variance);
}
Col euclidean(const Mat2D& sims, const Row& ref) { return ABC_B200::euclidean<Col>(sims, ref); }
// optional: replaces src/AbcUtil.cpp:378-390; two words of the caller's gsl_rng seed the Philox counters
Mat2D sample_predictive_priors(const gsl_rng* RNG, const size_t num_samples, const Col& weights, const Mat2D& parameter_prior,
                               const std::vector<const Parameter*>& pars, const Row& doubled_variance) {
    const uint64_t seed = ((uint64_t)gsl_rng_get(RNG) << 32) ^ (uint64_t)gsl_rng_get(RNG);
    return ABC_B200::sample_predictive_priors<Mat2D>(seed, num_samples, weights, parameter_prior, pars, doubled_variance);
}
}  // namespace ABC
#endif  // ABCB200_DROP_IN

#endif  // ABC_B200_HPP
