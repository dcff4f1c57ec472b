// TEST INFRASTRUCTURE (oracle/): C entry points around the reference's OWN functions, compiled together with the reference's
// unmodified sources (lib/PLS/src/pls.cpp, src/AbcUtil.cpp, taken where they lie under /root/reference) into
// oracle/_ref/libabcref.so by `make -C oracle ref`. Eigen and GSL are absent from this image; the sources are compiled against
// the stand-ins in oracle/shim/ (mini_eigen.hpp, gsl/gsl_stub.h — read their headers for what is and is not the reference's
// arithmetic). This library exists to PIN oracle/abc_oracle.cpp (the hand restatement the GPU parity tests use): every ref_*
// function below calls the reference function named in its comment and nothing else; tests/test_ref_pin.py compares the two and
// tests/golden/make_ref_fixtures.py stores the reference's outputs as fixtures that travel to the GPU box (this library does not
// need to: /root/reference does not exist there).
// The signatures mirror the orc_* entry points of abc_oracle.cpp (column-major doubles, long sizes) so that oracle/ref.py can
// reuse the oracle's ctypes wrappers.
#include <AbcSmc/AbcUtil.h>
#include <AbcSmc/Priors.h>
#include <PLS/pls.h>

#include <cstdint>
#include <memory>
#include <random>
#include <sstream>
#include <string>
#include <vector>

namespace {

Mat2D to_mat(const double* p, long n, long k) {
    Mat2D m(n, k);
    for (long j = 0; j < k; j++) for (long i = 0; i < n; i++) m(i, j) = p[i + j * n];
    return m;
}
Row to_row(const double* p, long k) { Row r(k); for (long j = 0; j < k; j++) r[j] = p[j]; return r; }
Col to_col(const double* p, long n) { Col c(n); for (long i = 0; i < n; i++) c[i] = p[i]; return c; }
template <class M> void from_mat(const M& m, double* out) {
    for (long j = 0; j < m.cols(); j++) for (long i = 0; i < m.rows(); i++) out[i + j * m.rows()] = m(i, j);
}
template <class V> void from_vec(const V& v, double* out) { for (long i = 0; i < v.size(); i++) out[i] = v[i]; }

struct RefModel {
    Mat2D X, Y;
    std::unique_ptr<PLS::Model> m;
    long A;
};

// prior kinds as in abc_oracle.cpp: 0 uniform [a, b], 1 discrete uniform [a, b], 2 Gaussian(mean a, sd b)
std::unique_ptr<ABC::Parameter> make_prior(int type, double a, double b) {
    switch (type) {
        case 0: return std::unique_ptr<ABC::Parameter>(new ABC::ContinuousUniformPrior("p", "p", a, b));
        case 1: return std::unique_ptr<ABC::Parameter>(new ABC::DiscreteUniformPrior("p", "p", (long)a, (long)b));
        default: return std::unique_ptr<ABC::Parameter>(new ABC::GaussianPrior("p", "p", a, b));
    }
}

// Model keeps P, W, R, Q, T private; print_state (pls.cpp:565-579) streams them. At 17 significant digits the text round-trips.
// Complex entries are streamed as "(re,im)".
std::vector<double> parse_state_block(const std::string& text, const std::string& label, const std::string& next_label) {
    const size_t b = text.find(label + ":\n");
    const size_t e = next_label.empty() ? text.size() : text.find(next_label + ":\n", b);
    std::vector<double> re;
    if (b == std::string::npos) return re;
    std::string body = text.substr(b + label.size() + 2, e - (b + label.size() + 2));
    size_t pos = 0;
    while ((pos = body.find('(', pos)) != std::string::npos) {
        const size_t comma = body.find(',', pos);
        re.push_back(std::stod(body.substr(pos + 1, comma - pos - 1)));
        pos = comma;
    }
    return re;   // row-major, as printed
}

}  // namespace

extern "C" {

// Eigen colwise().mean() as AbcUtil.cpp:412, 432 call it
void ref_colwise_mean(const double* X, long n, long k, double* out) { Row m = to_mat(X, n, k).colwise().mean(); from_vec(m, out); }
// PLS::colwise_stdev (pls.cpp:79-83)
void ref_colwise_stdev(const double* X, long n, long k, const double* mean, double* out) { Row s = PLS::colwise_stdev(to_mat(X, n, k), to_row(mean, k)); from_vec(s, out); }
// PLS::colwise_z_scores (pls.cpp:93-105 and :107-111)
void ref_colwise_z_scores(const double* X, long n, long k, const double* mean, const double* sd, double* Z) { from_mat(PLS::colwise_z_scores(to_mat(X, n, k), to_row(mean, k), to_row(sd, k)), Z); }
void ref_colwise_z_scores_auto(const double* X, long n, long k, double* Z) { from_mat(PLS::colwise_z_scores(to_mat(X, n, k)), Z); }
// PLS::z_scores (pls.cpp:89-91)
void ref_z_scores(const double* obs, const double* mean, const double* sd, long k, double* out) { Row z = PLS::z_scores(to_row(obs, k), to_row(mean, k), to_row(sd, k)); from_vec(z, out); }
// PLS::normalcdf (pls.cpp:152-160), PLS::wilcoxon (:190-211)
double ref_normalcdf(double z) { return PLS::normalcdf(z); }
double ref_wilcoxon(const double* e1, const double* e2, long n) { return PLS::wilcoxon(to_col(e1, n), to_col(e2, n)); }
// PLS::ordered (pls.h:58-69)
void ref_ordered(const double* v, long n, uint64_t* out) { auto o = PLS::ordered(to_col(v, n)); for (long i = 0; i < n; i++) out[i] = o[i]; }
// ABC::euclidean (AbcUtil.cpp:320-324)
void ref_euclidean(const double* S, long n, long k, const double* ref, double* out) { Col d = ABC::euclidean(to_mat(S, n, k), to_row(ref, k)); from_vec(d, out); }

// PLS::Model::Model (pls.cpp:340-359) -> plsr (:390-437)
void* ref_pls_fit(const double* X, const double* Y, long n, long K, long M, int method, long max_components) {
    RefModel* r = new RefModel{to_mat(X, n, K), to_mat(Y, n, M), nullptr, max_components};
    r->m.reset(new PLS::Model(r->X, r->Y, method == 0 ? PLS::KERNEL_TYPE1 : PLS::KERNEL_TYPE2, (size_t)max_components));
    return r;
}
void ref_pls_free(void* p) { delete (RefModel*)p; }
// which in {P, W, R, Q, T}: real parts of the private factors through Model::print_state (pls.cpp:565-579)
int ref_pls_get(void* p, char which, double* out) {
    RefModel* r = (RefModel*)p;
    std::ostringstream os; os.precision(17);
    r->m->print_state(os);
    const std::string text = os.str();
    const char* order[] = {"P", "W", "R", "Q", "T", "coefficients"};
    for (int i = 0; i < 5; i++) if (order[i][0] == which) {
        std::vector<double> v = parse_state_block(text, order[i], order[i + 1]);
        const long cols = r->A, rows = cols ? (long)v.size() / cols : 0;
        if (rows * cols != (long)v.size()) return -1;
        for (long a = 0; a < rows; a++) for (long b = 0; b < cols; b++) out[a + b * rows] = v[(size_t)(a * cols + b)];
        return (int)rows;
    }
    return -1;
}
// Model::scores (pls.cpp:439-442), real parts as AbcUtil.cpp:453-454 takes them
void ref_pls_scores(void* p, const double* Xn, long n, long comp, double* out) { RefModel* r = (RefModel*)p; from_mat(r->m->scores(to_mat(Xn, n, r->X.cols()), (size_t)comp).real(), out); }
// Model::coefficients (pls.cpp:444-447), real parts as :450 takes them
void ref_pls_coefficients(void* p, long comp, double* out) { from_mat(((RefModel*)p)->m->coefficients((size_t)comp).real(), out); }
// Model::fitted_values / residuals / SSE / explained_variance (pls.cpp:449-467)
void ref_pls_fitted_values(void* p, const double* Xn, long n, long comp, double* out) { RefModel* r = (RefModel*)p; from_mat(r->m->fitted_values(to_mat(Xn, n, r->X.cols()), (size_t)comp), out); }
void ref_pls_residuals(void* p, const double* Xn, const double* Yn, long n, long comp, double* out) { RefModel* r = (RefModel*)p; from_mat(r->m->residuals(to_mat(Xn, n, r->X.cols()), to_mat(Yn, n, r->Y.cols()), (size_t)comp), out); }
void ref_pls_SSE(void* p, const double* Xn, const double* Yn, long n, long comp, double* out) { RefModel* r = (RefModel*)p; Row s = r->m->SSE(to_mat(Xn, n, r->X.cols()), to_mat(Yn, n, r->Y.cols()), (size_t)comp); from_vec(s, out); }
void ref_pls_explained_variance(void* p, const double* Xn, const double* Yn, long n, long comp, double* out) { RefModel* r = (RefModel*)p; Row s = r->m->explained_variance(to_mat(Xn, n, r->X.cols()), to_mat(Yn, n, r->Y.cols()), (size_t)comp); from_vec(s, out); }
// Model::cv_NEW_DATA (pls.cpp:494-510), cv_LOO (:469-491), cv_LSO (:512-549)
void* ref_pls_cv_new_data(void* p, const double* Xn, const double* Yn, long n) { RefModel* r = (RefModel*)p; return new PLS::Residual(r->m->cv_NEW_DATA(to_mat(Xn, n, r->X.cols()), to_mat(Yn, n, r->Y.cols()))); }
void* ref_pls_cv_loo(void* p) { return new PLS::Residual(((RefModel*)p)->m->cv_LOO()); }
void* ref_pls_cv_lso_seeded(void* p, double test_fraction, long num_trials, uint32_t seed) { std::mt19937 rng(seed); return new PLS::Residual(((RefModel*)p)->m->cv_LSO(test_fraction, (size_t)num_trials, rng)); }
// the `full` vector after each PLS::rand_nchoosek call (pls.cpp:218-227) of a cv_LSO run with the same seed: the partitions the
// oracle's cv_LSO takes as an explicit input
void ref_lso_shuffles(uint32_t seed, long N, long test_size, long num_trials, uint64_t* out) {
    std::mt19937 rng(seed);
    std::vector<Eigen::Index> full((size_t)N), sample((size_t)(N - test_size)), complement((size_t)test_size);
    for (long i = 0; i < N; i++) full[(size_t)i] = i;
    for (long t = 0; t < num_trials; t++) {
        PLS::rand_nchoosek(rng, full, sample, complement);
        for (long i = 0; i < N; i++) out[t * N + i] = (uint64_t)full[(size_t)i];
    }
}
void ref_residual_free(void* r) { delete (PLS::Residual*)r; }
long ref_residual_rows(void* r) { auto e = ((PLS::Residual*)r)->errors(); return e.empty() ? 0 : (long)e[0].rows(); }
void ref_residual_errors(void* rp, double* out) { auto e = ((PLS::Residual*)rp)->errors(); size_t off = 0; for (auto& m : e) { from_mat(m, out + off); off += (size_t)m.size(); } }
// PLS::validation (pls.cpp:235-261), PLS::optimal_num_components (:265-289)
void ref_validation(void* rp, int out_type, double* out) { from_mat(PLS::validation(*(PLS::Residual*)rp, out_type == 0 ? PLS::RESS : PLS::MSE), out); }
void ref_optimal_num_components(void* rp, double alpha, uint64_t* out) { Colsz v = PLS::optimal_num_components(*(PLS::Residual*)rp, alpha); for (long i = 0; i < v.size(); i++) out[i] = v[i]; }

// ABC::particle_ranking_PLS (AbcUtil.cpp:423-458), ABC::particle_ranking_simple (:408-421): the order is all they return
void ref_particle_ranking_PLS(const double* met, const double* par, long N, long K, long P, const double* target, double training_fraction, uint64_t* order_out) {
    auto o = ABC::particle_ranking_PLS(to_mat(met, N, K), to_mat(par, N, P), to_row(target, K), training_fraction);
    for (long i = 0; i < N; i++) order_out[i] = o[(size_t)i];
}
void ref_particle_ranking_simple(const double* met, long N, long K, const double* target, uint64_t* order_out) {
    auto o = ABC::particle_ranking_simple(to_mat(met, N, K), Mat2D(), to_row(target, K));
    for (long i = 0; i < N; i++) order_out[i] = o[(size_t)i];
}
// ABC::calculate_doubled_variance (AbcUtil.cpp:528-537)
void ref_calculate_doubled_variance(const double* params, long n, long P, double* out) { Row v = ABC::calculate_doubled_variance(to_mat(params, n, P)); from_vec(v, out); }
// Prior::likelihood (Priors.h:54-56, 75-77, 101-103)
double ref_prior_likelihood(int type, double a, double b, double v) { return make_prior(type, a, b)->likelihood(v); }
// ABC::weight_predictive_prior, set 0 (AbcUtil.cpp:539-545) and set > 0 (:547-586)
void ref_weight_predictive_prior0(long n, long P, double* out) {
    std::vector<const ABC::Parameter*> none;
    Row w = ABC::weight_predictive_prior(none, Mat2D::Zero(n, P));
    from_vec(w, out);
}
void ref_weight_predictive_prior(const int* ptype, const double* pa, const double* pb, const double* params, long n_new, const double* prev_params, long n_old,
                                 const double* prev_w, const double* prev_dv, long P, double* out) {
    std::vector<std::unique_ptr<ABC::Parameter>> own;
    std::vector<const ABC::Parameter*> mpars;
    for (long p = 0; p < P; p++) { own.push_back(make_prior(ptype[p], pa[p], pb[p])); mpars.push_back(own.back().get()); }
    Row w = ABC::weight_predictive_prior(mpars, to_mat(params, n_new, P), to_mat(prev_params, n_old, P), to_row(prev_w, n_old), to_row(prev_dv, P));
    from_vec(w, out);
}
// AbcLog::filtering_report statistics: ABC::calculate_nrmse (AbcUtil.cpp:326-345), ABC::median (:46-61)
double ref_calculate_nrmse(const double* mets, long n, long k, const double* observed) { return ABC::calculate_nrmse(to_mat(mets, n, k), to_row(observed, k)); }
double ref_median(const double* v, long n) { return ABC::median(to_col(v, n)); }

// Next-set proposals. ABC::setup_mvn_sampler (AbcUtil.cpp:462-488) is deterministic: L (lower triangle incl.
// the diagonal; the upper triangle is zeroed here, GSL leaves the covariance in it). The samplers (AbcUtil.cpp:378-404,
// Priors.h:18-41) run on the stand-in's MT19937 stream: distributional checks only.
void ref_setup_mvn_sampler(const double* theta, long n_pp, long P, double* L_out) {
    gsl_matrix* L = ABC::setup_mvn_sampler(to_mat(theta, n_pp, P));
    for (long j = 0; j < P; j++) for (long i = 0; i < P; i++) L_out[i + j * P] = i >= j ? gsl_matrix_get(L, (size_t)i, (size_t)j) : 0.0;
    gsl_matrix_free(L);
}
static std::vector<const ABC::Parameter*> make_priors(const int* ptype, const double* pa, const double* pb, long P, std::vector<std::unique_ptr<ABC::Parameter>>& own) {
    std::vector<const ABC::Parameter*> mpars;
    for (long p = 0; p < P; p++) { own.push_back(make_prior(ptype[p], pa[p], pb[p])); mpars.push_back(own.back().get()); }
    return mpars;
}
void ref_sample_predictive_priors(uint32_t seed, long num_samples, const double* weights, const double* theta, long n_pp, long P, const int* ptype,
                                  const double* pa, const double* pb, const double* dv, double* out) {
    gsl_rng rng; rng.eng.seed(seed);
    std::vector<std::unique_ptr<ABC::Parameter>> own;
    const auto mpars = make_priors(ptype, pa, pb, P, own);
    from_mat(ABC::sample_predictive_priors(&rng, (size_t)num_samples, to_col(weights, n_pp), to_mat(theta, n_pp, P), mpars, to_row(dv, P)), out);
}
void ref_sample_mvn_predictive_priors(uint32_t seed, long num_samples, const double* weights, const double* theta, long n_pp, long P, const int* ptype,
                                      const double* pa, const double* pb, const double* L_colmajor, double* out) {
    gsl_rng rng; rng.eng.seed(seed);
    std::vector<std::unique_ptr<ABC::Parameter>> own;
    const auto mpars = make_priors(ptype, pa, pb, P, own);
    gsl_matrix* L = gsl_matrix_alloc((size_t)P, (size_t)P);
    for (long i = 0; i < P; i++) for (long j = 0; j < P; j++) gsl_matrix_set(L, (size_t)i, (size_t)j, L_colmajor[i + j * P]);
    from_mat(ABC::sample_mvn_predictive_priors(&rng, (size_t)num_samples, to_col(weights, n_pp), to_mat(theta, n_pp, P), mpars, L), out);
    gsl_matrix_free(L);
}

}  // extern "C"
