// ref_stub.hpp — stand-ins for the reference's types, just enough for abc_b200.hpp's ABCB200_DROP_IN / ABCB200_DROP_IN_PLS blocks
// to compile and run without Eigen and GSL (neither is in this image; the reference needs both):
//   float_type, Mat2D, Col, Row, Colsz    lib/PLS/include/PLS/pls.h:21-33 (Eigen::MatrixXd / VectorXd / RowVectorXd / Matrix<size_t,-1,1>)
//   Parameter                             include/AbcSmc/Parameter.h:29-85 (the virtuals the hot path calls) + a ContinuousUniformPrior (Priors.h:86-110)
//   gsl_rng, gsl_rng_get                  <gsl/gsl_rng.h>
// and the prototypes of the namespace-ABC functions the drop-in defines (include/AbcSmc/AbcUtil.h:103, 121-126, 146-172), restated.
// Column-major storage, data() / rows() / cols() / outerStride() / size() as Eigen has them. Row and Col are DISTINCT types, as in Eigen.
#ifndef ABCB200_REF_STUB_HPP
#define ABCB200_REF_STUB_HPP
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <ostream>
#include <vector>

typedef double float_type;

template <class S>
struct StubMat {
    StubMat() : r_(0), c_(0) {}
    StubMat(long r, long c) : r_(r), c_(c), d_((size_t)r * (size_t)c) {}
    S* data() { return d_.data(); }
    const S* data() const { return d_.data(); }
    long rows() const { return r_; }
    long cols() const { return c_; }
    long outerStride() const { return r_; }
    S& operator()(long i, long j) { return d_[(size_t)j * (size_t)r_ + (size_t)i]; }
    const S& operator()(long i, long j) const { return d_[(size_t)j * (size_t)r_ + (size_t)i]; }
    long r_, c_;
    std::vector<S> d_;
};
template <class S, int TAG>
struct StubVec {
    StubVec() {}
    explicit StubVec(long n) : d_((size_t)n) {}
    S* data() { return d_.data(); }
    const S* data() const { return d_.data(); }
    long size() const { return (long)d_.size(); }
    S& operator[](long i) { return d_[(size_t)i]; }
    const S& operator[](long i) const { return d_[(size_t)i]; }
    std::vector<S> d_;
};
typedef StubMat<double> Mat2D;
typedef StubVec<double, 0> Col;
typedef StubVec<double, 1> Row;
typedef StubVec<size_t, 2> Colsz;
template <class S> std::ostream& operator<<(std::ostream& os, const StubMat<S>& m) {
    for (long i = 0; i < m.rows(); i++) { for (long j = 0; j < m.cols(); j++) os << (j ? " " : "") << m(i, j); if (i + 1 < m.rows()) os << "\n"; }
    return os;
}
template <class S, int T> std::ostream& operator<<(std::ostream& os, const StubVec<S, T>& v) {
    for (long i = 0; i < v.size(); i++) os << (i ? " " : "") << v[i];
    return os;
}

struct gsl_rng { uint64_t state; };
inline unsigned long gsl_rng_get(const gsl_rng* r) {      // any 32-bit generator will do for the stand-in
    uint64_t& s = const_cast<gsl_rng*>(r)->state;
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    return (unsigned long)(s >> 32);
}

class Parameter {
  public:
    virtual ~Parameter() {}
    virtual float_type recast(const float_type pval) const = 0;
    virtual float_type likelihood(const float_type pval) const = 0;
    virtual float_type get_mean() const = 0;
    virtual float_type get_sd() const = 0;
    bool valid(const float_type pval) const { return likelihood(pval) != 0.0; }
};
class ContinuousUniformPrior : public Parameter {     // Priors.h:86-110
  public:
    ContinuousUniformPrior(double a, double b) : a_(a), b_(b) {}
    float_type recast(const float_type pval) const override { return pval; }
    float_type likelihood(const float_type v) const override { return (v >= a_ && v <= b_) ? 1.0 / (b_ - a_) : 0.0; }
    float_type get_mean() const override { return (a_ + b_) / 2.0; }
    float_type get_sd() const override { return (b_ - a_) / std::sqrt(12.0); }
    double a_, b_;
};

namespace ABC {
Col euclidean(const Mat2D& sims, const Row& ref);
Mat2D sample_predictive_priors(const gsl_rng* RNG, const size_t num_samples, const Col& weights, const Mat2D& parameter_prior,
                               const std::vector<const Parameter*>& pars, const Row& doubled_variance);
std::vector<size_t> particle_ranking_simple(const Mat2D& X_orig, const Mat2D& Y_orig, const Row& target_values);
std::vector<size_t> particle_ranking_PLS(const Mat2D& X_orig, const Mat2D& Y_orig, const Row& target_values, const float_type training_fraction);
Row weight_predictive_prior(const std::vector<const Parameter*>& mpars, const Mat2D& params);
Row weight_predictive_prior(const std::vector<const Parameter*>& mpars, const Mat2D& params, const Mat2D& prev_params, const Row& prev_weights,
                            const Row& prev_doubled_variance);
Row calculate_doubled_variance(const Mat2D& params);
}  // namespace ABC
#endif
