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
#include <random>
#include <fstream>
#include <cmath>
#include <limits>
#include <algorithm>
#include <vector>
#include <string>
#include <memory>
#include <functional>
#include <iomanip>

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
    // 0: exact distance ties in ascending particle index; 1: where libstdc++'s std::sort leaves them in PLS::ordered (pls.h:58-69)
    void set_tie_order(int mode) { check(abcb200_set_tie_order(h_, mode), "abcb200_set_tie_order"); }
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
#if (defined(ABCB200_DROP_IN) || defined(ABCB200_DROP_IN_PLS)) && !defined(ABCB200_TIES_BY_INDEX)
        // a drop-in returns what the reference returns, tie groups included (a host pass only when exact ties reach the output)
        abcb200_set_tie_order(h_, 1);
#endif
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

// One call per SMC set with set t-1 resident on the device (include/abcsmc_b200.h: abcb200_chain_*): what the body
// of AbcSmc::read_SMC_sets_from_database's set loop computes (src/AbcSmc.cpp:634-664) plus calculate_predictive_prior_weights
// (:1041-1066), without re-uploading or re-evaluating anything of the earlier sets. The numerator of the weights comes from the host's
// Parameter objects (one virtual call per particle and parameter, as src/AbcUtil.cpp:559-561).
template <class Mat2D, class Row>
class SmcChain {
  public:
    struct SetResult {
        std::vector<size_t> predictive_prior;      // _predictive_prior[t]: particle indices by rank
        Row weights, doubled_variance;             // _weights[t], _doubled_variance[t]
        double nrmse;                              // AbcLog::filtering_report: ABC::calculate_nrmse(posterior_mets, observed)
        Row mean_par, mean_met, median_par, median_met;
        int n_components;                          // PLS components used (0 for the SIMPLE filter)
    };
    explicit SmcChain(int n_params) : P_(n_params), h_(nullptr) {
        auto& c = abcb200::Context::instance();
        c.check(abcb200_chain_create(c.handle(), n_params, &h_), "SmcChain");
    }
    ~SmcChain() { abcb200_chain_destroy(h_); }
    SmcChain(const SmcChain&) = delete;
    SmcChain& operator=(const SmcChain&) = delete;
    int sets() const { return abcb200_chain_sets(h_); }

    // use_pls: FILTER::PLS (particle_ranking_PLS) or FILTER::SIMPLE (particle_ranking_simple)
    template <class Parameter>
    SetResult process_set(const Mat2D& metrics, const Mat2D& params, const Row& observed, const std::vector<const Parameter*>& mpars, bool use_pls,
                          double training_fraction, size_t next_pred_prior_size) {
        auto& c = abcb200::Context::instance();
        const int64_t N = (int64_t)metrics.rows();
        const int K = (int)metrics.cols();
        std::vector<double> numer;
        if (sets() > 0) {       // set 0 needs no numerator (uniform 1 / n, AbcUtil.cpp:539-545)
            numer.assign((size_t)N, 1.0);
            for (int p = 0; p < P_; p++) {
                const double* col = params.data() + (int64_t)p * abcb200::ld(params);
                for (int64_t i = 0; i < N; i++) numer[(size_t)i] *= mpars[(size_t)p]->likelihood(col[i]);
            }
        }
        const size_t n = (next_pred_prior_size == 0 || next_pred_prior_size > (size_t)N) ? (size_t)N : next_pred_prior_size;
        std::vector<uint64_t> order(n);
        std::vector<double> rep((size_t)(1 + 2 * (P_ + K)));
        SetResult r{std::vector<size_t>(), Row((long)n), Row(P_), 0.0, Row(P_), Row(K), Row(P_), Row(K), 0};
        c.check(abcb200_chain_process_set(h_, metrics.data(), abcb200::ld(metrics), params.data(), abcb200::ld(params), N, K, observed.data(), use_pls ? 0 : 1,
                                          training_fraction, ABCB200_KERNEL_TYPE1, (int64_t)n, nullptr, nullptr, nullptr, numer.empty() ? nullptr : numer.data(),
                                          order.data(), r.weights.data(), r.doubled_variance.data(), rep.data(), &r.n_components),
                "SmcChain::process_set");
        r.predictive_prior.assign(order.begin(), order.end());
        r.nrmse = rep[0];
        std::copy(rep.begin() + 1, rep.begin() + 1 + P_, r.mean_par.data());
        std::copy(rep.begin() + 1 + P_, rep.begin() + 1 + P_ + K, r.mean_met.data());
        std::copy(rep.begin() + 1 + P_ + K, rep.begin() + 1 + 2 * P_ + K, r.median_par.data());
        std::copy(rep.begin() + 1 + 2 * P_ + K, rep.end(), r.median_met.data());
        return r;
    }

  private:
    int P_;
    abcb200_chain* h_;
};

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

// class PLS::Residual (pls.h:44-53) with the reference's public surface: errors() and method(). The reference's object IS
// the M matrices n x A; here the cross-validation that produced it has already streamed PRESS and the component counts on
// the device, and the error cube is only materialised (host memory, the reference's layout) when errors() is asked for,
// or when optimal_num_components() is called with an ALPHA other than the one the streamed selection used.
template <class Mat2D>
class Residual {
  public:
    const std::vector<Mat2D> errors() const {                           // pls.h:51 (returns a copy, as the reference)
        if (!s_->have_cube) { s_->cube = s_->materialise(); s_->have_cube = true; }
        return s_->cube;
    }
    const std::string method() const { return s_->label; }              // pls.h:52
    // streamed results (not in the reference's class; what validation() / optimal_num_components() below return)
    const Mat2D& ress() const { return s_->ress; }
    int64_t rows() const { return s_->n; }
    const std::vector<size_t>& streamed_num_components() const { return s_->ncomp; }
    double streamed_alpha() const { return s_->alpha; }

    struct State {
        std::string label;
        int64_t n;
        Mat2D ress;                                  // M x A, sum of squared errors (RESS)
        std::vector<size_t> ncomp;
        double alpha;
        std::function<std::vector<Mat2D>()> materialise;
        std::vector<Mat2D> cube;
        bool have_cube;
    };
    explicit Residual(std::shared_ptr<State> s) : s_(std::move(s)) {}

  private:
    std::shared_ptr<State> s_;
};

// Mat2D PLS::validation(residual, out_type) (pls.cpp:235-261): M x A; MSE = RESS / n_obs
template <class Mat2D>
Mat2D validation(const Residual<Mat2D>& residual, const VALIDATION_OUTPUT out_type) {
    Mat2D out(residual.ress());
    if (out_type == MSE) {
        const double n = (double)residual.rows();
        for (long c = 0; c < (long)out.cols(); c++) for (long r = 0; r < (long)out.rows(); r++) out.data()[(size_t)c * (size_t)out.outerStride() + (size_t)r] /= n;
    }
    return out;
}
// Colsz PLS::optimal_num_components(residual, ALPHA) (pls.cpp:265-289) as std::vector<size_t>
template <class Mat2D>
std::vector<size_t> optimal_num_components(const Residual<Mat2D>& residual, const double ALPHA = 0.1) {
    if (ALPHA == residual.streamed_alpha()) return residual.streamed_num_components();
    const std::vector<Mat2D> ev = residual.errors();
    const int M = (int)ev.size();
    const int64_t n = (int64_t)ev[0].rows();
    const int A = (int)ev[0].cols();
    std::vector<double> flat((size_t)M * (size_t)n * (size_t)A);
    for (int y = 0; y < M; y++)
        for (int c = 0; c < A; c++)
            std::copy(ev[(size_t)y].data() + (size_t)c * (size_t)ev[(size_t)y].outerStride(), ev[(size_t)y].data() + (size_t)c * (size_t)ev[(size_t)y].outerStride() + n,
                      flat.begin() + (std::ptrdiff_t)(((size_t)y * A + c) * (size_t)n));
    std::vector<int32_t> nc((size_t)M);
    auto& c = abcb200::Context::instance();
    c.check(abcb200_residual_select(c.handle(), flat.data(), n, M, A, ABCB200_RESS, ALPHA, nullptr, nc.data()), "optimal_num_components");
    return std::vector<size_t>(nc.begin(), nc.end());
}
// void PLS::print_validation(residual, out_type, os) (pls.cpp:291-305); needs operator<<(ostream, Mat2D) as Eigen has
template <class Mat2D>
void print_validation(const Residual<Mat2D>& residual, const VALIDATION_OUTPUT out_type, std::ostream& os = std::cerr) {
    os << residual.method() << " Validation:" << std::endl;
    Mat2D em = validation(residual, out_type);
    switch (out_type) {
        case MSE:
            os << "RMSE ";
            for (long c = 0; c < (long)em.cols(); c++) for (long r = 0; r < (long)em.rows(); r++) { double& v = em.data()[(size_t)c * (size_t)em.outerStride() + (size_t)r]; v = std::sqrt(v); }
            break;
        case RESS: os << "PRESS "; break;
        default: os << "UNKNOWN ";
    }
    os << " Matrix (rows = Y variable; cols = # of components):" << std::endl << em << std::endl;
    os << "Optimal number of components (by Y variable):\t";
    for (size_t v : optimal_num_components(residual)) os << v << "\n";
    os << std::endl;
}

// struct PLS::Model (pls.h:184-266). Real parts only: every consumer of the reference's complex factors takes .real().
// Value semantics as in the reference: copies share the fitted factors on the device (they are immutable after the
// constructor) and the host copies _X, _Y the reference keeps for cv_LOO / cv_LSO (pls.cpp:344).
template <class Mat2D, class Row>
struct Model {
    Model(const Mat2D& X, const Mat2D& Y, const METHOD& algorithm, const size_t& max_components)
        : K_((int)X.cols()), M_((int)Y.cols()), A_((int)max_components), N_((int64_t)X.rows()), method_(algorithm),
          X_(std::make_shared<const Mat2D>(X)), Y_(std::make_shared<const Mat2D>(Y)) {
        auto& c = abcb200::Context::instance();
        abcb200_pls* h = nullptr;
        c.check(abcb200_pls_fit(c.handle(), X.data(), abcb200::ld(X), Y.data(), abcb200::ld(Y), N_, K_, M_, (int)algorithm, A_, &h), "PLS::Model");
        h_ = std::shared_ptr<abcb200_pls>(h, [](abcb200_pls* p) { abcb200_pls_free(p); });
    }
    Model(const Mat2D& X, const Mat2D& Y, const METHOD& algorithm = KERNEL_TYPE1) : Model(X, Y, algorithm, (size_t)X.cols()) {}

    const Mat2D scores(const Mat2D& X_new, const size_t comp) const {
        Mat2D out(X_new.rows(), comp);
        ck(abcb200_pls_scores(h_.get(), X_new.data(), abcb200::ld(X_new), (int64_t)X_new.rows(), (int)comp, out.data()), "Model::scores");
        return out;
    }
    const Mat2D scores(const Mat2D& X_new) const { return scores(X_new, (size_t)A_); }
    const Mat2D loadingsX(const size_t comp) const { return factor('P', K_, comp); }   // declared, never defined, in the reference (pls.h:207-211)
    const Mat2D loadingsX() const { return loadingsX((size_t)A_); }
    const Mat2D loadingsY(const size_t comp) const { return factor('Q', M_, comp); }
    const Mat2D loadingsY() const { return loadingsY((size_t)A_); }
    const Mat2D coefficients(const size_t comp) const {
        Mat2D out(K_, M_);
        ck(abcb200_pls_coefficients(h_.get(), (int)comp, out.data()), "Model::coefficients");
        return out;
    }
    const Mat2D coefficients() const { return coefficients((size_t)A_); }
    const Mat2D fitted_values(const Mat2D& X, const size_t comp) const {
        Mat2D out(X.rows(), M_);
        ck(abcb200_pls_fitted_values(h_.get(), X.data(), abcb200::ld(X), (int64_t)X.rows(), (int)comp, out.data()), "Model::fitted_values");
        return out;
    }
    const Mat2D fitted_values(const Mat2D& X) const { return fitted_values(X, (size_t)A_); }
    const Mat2D residuals(const Mat2D& X, const Mat2D& Y, const size_t comp) const {
        Mat2D out(X.rows(), M_);
        ck(abcb200_pls_residuals(h_.get(), X.data(), abcb200::ld(X), Y.data(), abcb200::ld(Y), (int64_t)X.rows(), (int)comp, out.data()), "Model::residuals");
        return out;
    }
    const Mat2D residuals(const Mat2D& X, const Mat2D& Y) const { return residuals(X, Y, (size_t)A_); }
    const Row SSE(const Mat2D& X, const Mat2D& Y, const size_t comp) const {
        Row out(M_);
        ck(abcb200_pls_sse(h_.get(), X.data(), abcb200::ld(X), Y.data(), abcb200::ld(Y), (int64_t)X.rows(), (int)comp, out.data()), "Model::SSE");
        return out;
    }
    const Row SSE(const Mat2D& X, const Mat2D& Y) const { return SSE(X, Y, (size_t)A_); }
    const Row explained_variance(const Mat2D& X, const Mat2D& Y, const size_t comp) const {        // pls.cpp:461-467
        Row out(M_);
        ck(abcb200_pls_explained_variance(h_.get(), X.data(), abcb200::ld(X), Y.data(), abcb200::ld(Y), (int64_t)X.rows(), (int)comp, out.data()), "Model::explained_variance");
        return out;
    }
    const Row explained_variance(const Mat2D& X, const Mat2D& Y) const { return explained_variance(X, Y, (size_t)A_); }

    // ---- cross-validation with the reference's signatures (pls.h:236-238): a Residual comes back -----------------------------
    // Residual Model::cv_NEW_DATA(X, Y) const (pls.cpp:494-510): PRESS and the component counts are streamed at once; the error
    // cube, if asked for, is built the way the reference builds it: column c of Ev[y] = column y of residuals(X, Y, c + 1).
    Residual<Mat2D> cv_NEW_DATA(const Mat2D& X, const Mat2D& Y) const {
        auto st = new_state("NEW DATA", (int64_t)X.rows());
        std::vector<int32_t> nc((size_t)M_);
        ck(abcb200_pls_cv_new_data(h_.get(), X.data(), abcb200::ld(X), Y.data(), abcb200::ld(Y), (int64_t)X.rows(), ABCB200_RESS, 0.1, st->ress.data(), nc.data()), "Model::cv_NEW_DATA");
        st->ncomp.assign(nc.begin(), nc.end());
        const Model self(*this);
        auto Xc = std::make_shared<const Mat2D>(X), Yc = std::make_shared<const Mat2D>(Y);
        st->materialise = [self, Xc, Yc]() {
            std::vector<Mat2D> ev((size_t)self.M_, Mat2D(Xc->rows(), self.A_));
            for (int c = 0; c < self.A_; c++) {
                const Mat2D res = self.residuals(*Xc, *Yc, (size_t)(c + 1));
                for (int y = 0; y < self.M_; y++)
                    std::copy(res.data() + (size_t)y * (size_t)res.outerStride(), res.data() + (size_t)y * (size_t)res.outerStride() + (size_t)res.rows(),
                              ev[(size_t)y].data() + (size_t)c * (size_t)ev[(size_t)y].outerStride());
            }
            return ev;
        };
        return Residual<Mat2D>(st);
    }
    // Residual Model::cv_LOO() const (pls.cpp:469-491) on the rows the model was built from; label "LOO"
    Residual<Mat2D> cv_LOO() const {
        auto st = new_state("LOO", N_);
        std::vector<int32_t> nc((size_t)M_);
        auto& c = abcb200::Context::instance();
        ck(abcb200_pls_cv_loo(c.handle(), X_->data(), abcb200::ld(*X_), Y_->data(), abcb200::ld(*Y_), N_, K_, M_, A_, ABCB200_RESS, 0.1, nullptr, st->ress.data(), nc.data()), "Model::cv_LOO");
        st->ncomp.assign(nc.begin(), nc.end());
        const Model self(*this);
        st->materialise = [self]() {
            std::vector<double> flat((size_t)self.M_ * (size_t)self.N_ * (size_t)self.A_);
            auto& cc = abcb200::Context::instance();
            self.ck(abcb200_pls_cv_loo(cc.handle(), self.X_->data(), abcb200::ld(*self.X_), self.Y_->data(), abcb200::ld(*self.Y_), self.N_, self.K_, self.M_, self.A_,
                                       ABCB200_RESS, 0.1, flat.data(), nullptr, nullptr), "Model::cv_LOO (errors)");
            return self.unflatten(flat, self.N_);
        };
        return Residual<Mat2D>(st);
    }
    // Residual Model::cv_LSO(test_fraction, num_trials, rng) const (pls.cpp:512-549); the partitions are drawn exactly as
    // PLS::rand_nchoosek does (std::shuffle of the running index vector on the caller's generator, pls.cpp:217-227); label "LSO"
    template <class RNG>
    Residual<Mat2D> cv_LSO(const double test_fraction, const size_t num_trials, RNG& rng) const {
        const size_t N = (size_t)N_;
        const size_t test_size = (size_t)(test_fraction * (double)N + 0.5);
        auto shuffles = std::make_shared<std::vector<uint64_t>>(N * num_trials);
        std::vector<uint64_t> full(N);
        for (size_t i = 0; i < N; i++) full[i] = i;
        for (size_t t = 0; t < num_trials; t++) {
            std::shuffle(full.begin(), full.end(), rng);
            std::copy(full.begin(), full.end(), shuffles->begin() + (std::ptrdiff_t)(t * N));
        }
        const int64_t n_err = (int64_t)(test_size * num_trials);
        auto st = new_state("LSO", n_err);
        std::vector<int32_t> nc((size_t)M_);
        auto& c = abcb200::Context::instance();
        ck(abcb200_pls_cv_lso(c.handle(), X_->data(), abcb200::ld(*X_), Y_->data(), abcb200::ld(*Y_), N_, K_, M_, A_, (int)method_, shuffles->data(), (int64_t)test_size,
                              (int64_t)num_trials, ABCB200_RESS, 0.1, nullptr, st->ress.data(), nc.data()), "Model::cv_LSO");
        st->ncomp.assign(nc.begin(), nc.end());
        const Model self(*this);
        st->materialise = [self, shuffles, test_size, num_trials, n_err]() {
            std::vector<double> flat((size_t)self.M_ * (size_t)n_err * (size_t)self.A_);
            auto& cc = abcb200::Context::instance();
            self.ck(abcb200_pls_cv_lso(cc.handle(), self.X_->data(), abcb200::ld(*self.X_), self.Y_->data(), abcb200::ld(*self.Y_), self.N_, self.K_, self.M_, self.A_,
                                       (int)self.method_, shuffles->data(), (int64_t)test_size, (int64_t)num_trials, ABCB200_RESS, 0.1, flat.data(), nullptr, nullptr),
                    "Model::cv_LSO (errors)");
            return self.unflatten(flat, n_err);
        };
        return Residual<Mat2D>(st);
    }

    // ---- the same cross-validations, results only (no Residual object): validation matrix + component counts -------------------
    Validation<Mat2D> cv_NEW_DATA_streamed(const Mat2D& X, const Mat2D& Y, const VALIDATION_OUTPUT out_type = RESS, const double ALPHA = 0.1) const {
        Validation<Mat2D> v{Mat2D(M_, A_), std::vector<size_t>((size_t)M_)};
        std::vector<int32_t> nc((size_t)M_);
        ck(abcb200_pls_cv_new_data(h_.get(), X.data(), abcb200::ld(X), Y.data(), abcb200::ld(Y), (int64_t)X.rows(), (int)out_type, ALPHA, v.press.data(), nc.data()), "Model::cv_NEW_DATA");
        for (int y = 0; y < M_; y++) v.num_components[(size_t)y] = (size_t)nc[(size_t)y];
        return v;
    }
    // cv_LOO on explicitly passed matrices (any X, Y of the model's shape)
    Validation<Mat2D> cv_LOO(const Mat2D& X, const Mat2D& Y, const VALIDATION_OUTPUT out_type = RESS, const double ALPHA = 0.1) const {
        Validation<Mat2D> v{Mat2D(M_, A_), std::vector<size_t>((size_t)M_)};
        std::vector<int32_t> nc((size_t)M_);
        ck(abcb200_pls_cv_loo(abcb200::Context::instance().handle(), X.data(), abcb200::ld(X), Y.data(), abcb200::ld(Y), (int64_t)X.rows(), K_, M_, A_,
                              (int)out_type, ALPHA, nullptr, v.press.data(), nc.data()), "Model::cv_LOO");
        for (int y = 0; y < M_; y++) v.num_components[(size_t)y] = (size_t)nc[(size_t)y];
        return v;
    }
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

    // ---- output methods (pls.cpp:551-582); need operator<<(ostream, Mat2D / Row) as Eigen has ------------------------------------
    void print_explained_variance(const Mat2D& X, const Mat2D& Y, std::ostream& os = std::cerr) const {
        const int wd = (int)std::ceil(std::log10((double)A_));
        for (size_t ncomp = 1; ncomp <= (size_t)A_; ncomp++) {
            os << std::setw(wd) << ncomp << " components explained variance: ";
            os << explained_variance(X, Y, ncomp);
            os << "  - SSE: " << SSE(X, Y, ncomp) << std::endl;
        }
    }
    void print_state(std::ostream& os = std::cerr) const {
        os << "P:" << std::endl << factor('P', K_, (size_t)A_) << std::endl << "W:" << std::endl << factor('W', K_, (size_t)A_) << std::endl
           << "R:" << std::endl << factor('R', K_, (size_t)A_) << std::endl << "Q:" << std::endl << factor('Q', M_, (size_t)A_) << std::endl << "T:" << std::endl;
        if (method_ == KERNEL_TYPE1) os << factor('T', (int)N_, (size_t)A_);          // KERNEL_TYPE2 never forms T (pls.cpp:422-424)
        os << std::endl << "coefficients:" << std::endl << coefficients() << std::endl;
    }

  private:
    void ck(int rc, const char* where) const { abcb200::Context::instance().check(rc, where); }
    Mat2D factor(char which, int rows, size_t comp) const {
        Mat2D full(rows, A_);
        ck(abcb200_pls_get(h_.get(), which, full.data()), "Model factor");
        Mat2D out(rows, comp);
        for (size_t a = 0; a < comp; a++) for (int r = 0; r < rows; r++) out.data()[a * (size_t)out.outerStride() + r] = full.data()[a * (size_t)full.outerStride() + r];
        return out;
    }
    std::shared_ptr<typename Residual<Mat2D>::State> new_state(const char* label, int64_t n) const {
        auto st = std::make_shared<typename Residual<Mat2D>::State>(typename Residual<Mat2D>::State{label, n, Mat2D(M_, A_), {}, 0.1, nullptr, {}, false});
        return st;
    }
    // errors[(y * A + c) * n + i] (the C ABI's cube layout) -> M matrices n x A (Residual::errors())
    std::vector<Mat2D> unflatten(const std::vector<double>& flat, int64_t n) const {
        std::vector<Mat2D> ev((size_t)M_, Mat2D((long)n, A_));
        for (int y = 0; y < M_; y++)
            for (int c = 0; c < A_; c++)
                std::copy(flat.begin() + (std::ptrdiff_t)(((size_t)y * A_ + c) * (size_t)n), flat.begin() + (std::ptrdiff_t)(((size_t)y * A_ + c + 1) * (size_t)n),
                          ev[(size_t)y].data() + (size_t)c * (size_t)ev[(size_t)y].outerStride());
        return ev;
    }
    int K_, M_, A_;
    int64_t N_;
    METHOD method_;
    std::shared_ptr<const Mat2D> X_, Y_;          // _X, _Y of the reference (pls.h:251)
    std::shared_ptr<abcb200_pls> h_;
};

}  // namespace PLS_B200

#ifdef ABCB200_DROP_IN
// Definitions (external linkage, include in exactly one .cpp) with the reference's exact names and types; compile
// this in place of the bodies in src/AbcUtil.cpp (INTEGRATION.md §2 lists them).
// Requires the reference's headers first (Eigen typedefs Mat2D / Row / Col / float_type of lib/PLS/include/PLS/pls.h:15-27
// and Parameter). Compiled and run against stand-ins of those typedefs by tests/cpp/dropin_test.cpp.
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
#ifdef ABCB200_DROP_IN_SAMPLER
// Opt-in: replaces the body at src/AbcUtil.cpp:378-390 — remove that one too when this macro is set.
// NOT the reference's random stream: two words of the caller's gsl_rng seed Philox counters (distributional parity only).
Mat2D sample_predictive_priors(const gsl_rng* RNG, const size_t num_samples, const Col& weights, const Mat2D& parameter_prior,
                               const std::vector<const Parameter*>& pars, const Row& doubled_variance) {
    const uint64_t seed = ((uint64_t)gsl_rng_get(RNG) << 32) ^ (uint64_t)gsl_rng_get(RNG);
    return ABC_B200::sample_predictive_priors<Mat2D>(seed, num_samples, weights, parameter_prior, pars, doubled_variance);
}
#endif  // ABCB200_DROP_IN_SAMPLER
}  // namespace ABC
#endif  // ABCB200_DROP_IN

#ifdef ABCB200_DROP_IN_PLS
// namespace PLS with the reference's names (lib/PLS/include/PLS/pls.h:58-266) on top of PLS_B200, for users of the PLS
// library itself (lib/PLS/src/main.cpp:19-41 compiles against this block unchanged apart from its include line).
// Requires the typedefs Mat2D / Row / Col / Colsz / float_type first. Real factors: scores / loadings / coefficients return
// Mat2D where the reference returns Mat2Dc with zero imaginary parts.
namespace PLS {
typedef PLS_B200::Residual<Mat2D> Residual;
typedef PLS_B200::Model<Mat2D, Row> Model;
typedef PLS_B200::METHOD METHOD;
typedef PLS_B200::VALIDATION_OUTPUT VALIDATION_OUTPUT;
using PLS_B200::KERNEL_TYPE1; using PLS_B200::KERNEL_TYPE2; using PLS_B200::RESS; using PLS_B200::MSE;
template <typename T> std::vector<size_t> ordered(const T& v) { return PLS_B200::ordered(v); }
inline Row colwise_stdev(const Mat2D& mat) { return PLS_B200::colwise_stdev<Row>(mat); }
inline Mat2D colwise_z_scores(const Mat2D& mat) { return PLS_B200::colwise_z_scores(mat); }
inline Mat2D colwise_z_scores(const Mat2D& mat, const Row& mean, const Row& stdev) { return PLS_B200::colwise_z_scores(mat, mean, stdev); }
inline float_type wilcoxon(const Col& err_1, const Col& err_2) { return (float_type)PLS_B200::wilcoxon(err_1, err_2); }
inline Mat2D validation(const Residual& residual, const VALIDATION_OUTPUT out_type) { return PLS_B200::validation(residual, out_type); }
inline Colsz optimal_num_components(const Residual& residual, const float_type ALPHA = 0.1) {
    const std::vector<size_t> v = PLS_B200::optimal_num_components(residual, (double)ALPHA);
    Colsz out((long)v.size());
    for (size_t i = 0; i < v.size(); i++) out.data()[i] = v[i];
    return out;
}
inline void print_validation(const Residual& residual, const VALIDATION_OUTPUT out_type, std::ostream& os = std::cerr) { PLS_B200::print_validation(residual, out_type, os); }
// The small host-side helpers of pls.h:71-125, so that a program written against the reference's header (lib/PLS/src/main.cpp)
// compiles unchanged. They are O(N K) conveniences outside the AbcSmc path and run on the host.
template <typename EIGENTYPE> std::vector<float_type> to_cvector(const EIGENTYPE& data) { return std::vector<float_type>(data.data(), data.data() + data.size()); }
template <typename EIGENTYPE> inline EIGENTYPE to_evector(const std::vector<float_type>& data) {
    EIGENTYPE v((long)data.size());
    for (size_t i = 0; i < data.size(); i++) v.data()[i] = data[i];
    return v;
}
inline std::vector<std::string> split(const std::string& s, const char separator = ',') {      // pls.cpp:23-34: k separators -> k + 1 fields
    std::vector<std::string> fields;
    size_t from = 0;
    for (size_t at = s.find(separator); at != std::string::npos; at = s.find(separator, from)) { fields.push_back(s.substr(from, at - from)); from = at + 1; }
    fields.push_back(s.substr(from));
    return fields;
}
inline Mat2D read_matrix_file(const std::string& filename, const char separator = ',') {       // pls.cpp:37-67: no header, one row per line
    std::ifstream in(filename);
    std::vector<std::vector<float_type>> rows;
    for (std::string line; in.is_open() && std::getline(in, line);) {
        const std::vector<std::string> fields = split(line, separator);
        std::vector<float_type> row;
        for (const std::string& f : fields) row.push_back(std::stod(f));
        if (!rows.empty() && rows[0].size() != row.size()) {
            std::cerr << "Error: row " << rows.size() << " has " << row.size() << " columns, but previous row(s) have " << rows[0].size() << " columns." << std::endl;
            exit(1);
        }
        rows.push_back(row);
    }
    Mat2D X((long)rows.size(), rows.empty() ? 0 : (long)rows[0].size());
    for (size_t i = 0; i < rows.size(); i++) for (size_t j = 0; j < rows[i].size(); j++) X.data()[j * (size_t)X.outerStride() + i] = rows[i][j];
    return X;
}
inline Row SST(const Mat2D& mat, const Row& means) {                                           // pls.cpp:69-73
    Row out((long)mat.cols());
    for (long j = 0; j < (long)mat.cols(); j++) {
        double s = 0;
        if (mat.rows() >= 2) for (long i = 0; i < (long)mat.rows(); i++) { const double d = mat.data()[(size_t)j * (size_t)mat.outerStride() + (size_t)i] - means.data()[j]; s += d * d; }
        out.data()[j] = s;
    }
    return out;
}
inline Row colwise_means_(const Mat2D& mat) {
    Row m((long)mat.cols());
    for (long j = 0; j < (long)mat.cols(); j++) { double s = 0; for (long i = 0; i < (long)mat.rows(); i++) s += mat.data()[(size_t)j * (size_t)mat.outerStride() + (size_t)i]; m.data()[j] = s / (double)mat.rows(); }
    return m;
}
inline Row SST(const Mat2D& mat) { return SST(mat, colwise_means_(mat)); }                     // pls.cpp:75-77
inline Row colwise_stdev(const Mat2D& mat, const Row& means) {                                 // pls.cpp:79-83
    Row out = SST(mat, means);
    for (long j = 0; j < (long)mat.cols(); j++) out.data()[j] = std::sqrt(out.data()[j] / ((double)mat.rows() - 1));
    return out;
}
inline Row z_scores(const Row& obs, const Row& mean, const Row& stdev) {                       // pls.cpp:89-91 (no zero guard)
    Row out((long)obs.size());
    for (long j = 0; j < (long)obs.size(); j++) out.data()[j] = (obs.data()[j] - mean.data()[j]) / stdev.data()[j];
    return out;
}
inline float_type normalcdf(const float_type z) {                                              // pls.cpp:152-160: the 4-term A&S rational form
    const double a = std::fabs((double)z), poly = 1 + 0.196854 * a + 0.115194 * a * a + 0.000344 * a * a * a + 0.019527 * a * a * a * a;
    const double p = 0.5 / std::pow(poly, 4);
    return z < 0 ? p : 1.0 - p;
}
template <class RNG, class IDX>
inline void rand_nchoosek(RNG& rng, std::vector<IDX>& full, std::vector<IDX>& sample, std::vector<IDX>& complement) {     // pls.cpp:218-227
    std::shuffle(full.begin(), full.end(), rng);
    std::copy(full.begin(), full.begin() + (std::ptrdiff_t)sample.size(), sample.begin());
    std::copy(full.begin() + (std::ptrdiff_t)sample.size(), full.end(), complement.begin());
}
}  // namespace PLS
#endif  // ABCB200_DROP_IN_PLS

#endif  // ABC_B200_HPP
