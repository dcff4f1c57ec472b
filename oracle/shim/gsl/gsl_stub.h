// TEST INFRASTRUCTURE (oracle/): declarations of the GSL entry points the reference's src/AbcUtil.cpp and headers name, so that
// the UNMODIFIED reference source compiles in this image (GSL is absent, there is no network). Implemented for real, from GSL's
// documented behaviour: gsl_vector / gsl_matrix storage and gsl_ran_gaussian_pdf (randist/gauss.c: u = x / fabs(sigma);
// p = (1 / (sqrt(2 pi) fabs(sigma))) exp(-u u / 2)) — the only GSL arithmetic on the hot path (AbcUtil.cpp:574, Priors.h:54).
// Also implemented, from GSL's documented algorithms, for the callers right after the path (AbcUtil.cpp:111-158,
// 366-404, 462-488; Priors.h:18-41): gsl_ran_multivariate_gaussian_vcov (sample covariance, n - 1), gsl_linalg_cholesky_decomp1
// (lower factor in place, upper triangle untouched) — deterministic, so ABC::setup_mvn_sampler is pinned exactly — and a random
// stream: gsl_rng as MT19937 (GSL's default generator is the same algorithm; uniform = get / 2^32, uniform_int by rejection on
// get / (range / n)), gsl_ran_gaussian (polar Box-Muller, as randist/gauss.c), gsl_ran_multivariate_gaussian (mu + L z, z drawn in
// index order), gsl_ran_discrete (GSL builds Walker alias tables; here the cumulative weights are inverted with the same single
// uniform per draw: the same distribution, not the same stream). Sampling parity is distributional by construction.
// Special functions and minimisers (logistic regression, outside the path) abort with a message when called.
#pragma once
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <random>

#define GSL_SUCCESS 0
#define GSL_CONTINUE (-2)

extern "C++" {
struct gsl_rng { std::mt19937 eng; };
struct gsl_vector { size_t size; double* data; };
struct gsl_matrix { size_t size1, size2; double* data; };     // row-major like GSL
struct gsl_ran_discrete_t { size_t K; double* cdf; };
struct gsl_multimin_function { double (*f)(const gsl_vector*, void*); size_t n; void* params; };
struct gsl_multimin_fminimizer_type { const char* name; };
struct gsl_multimin_fminimizer { gsl_vector* x; double fval; double size; };

[[noreturn]] inline void gsl_stub_unavailable(const char* what) {
    std::fprintf(stderr, "oracle/shim/gsl: %s is not available in the GSL stand-in (outside the pinned path)\n", what);
    std::abort();
}

inline gsl_vector* gsl_vector_alloc(size_t n) { gsl_vector* v = new gsl_vector; v->size = n; v->data = new double[n ? n : 1](); return v; }
inline void gsl_vector_free(gsl_vector* v) { if (v) { delete[] v->data; delete v; } }
inline double gsl_vector_get(const gsl_vector* v, size_t i) { return v->data[i]; }
inline void gsl_vector_set(gsl_vector* v, size_t i, double x) { v->data[i] = x; }
inline void gsl_vector_set_all(gsl_vector* v, double x) { for (size_t i = 0; i < v->size; i++) v->data[i] = x; }
inline gsl_matrix* gsl_matrix_alloc(size_t r, size_t c) { gsl_matrix* m = new gsl_matrix; m->size1 = r; m->size2 = c; m->data = new double[(r * c != 0) ? r * c : 1](); return m; }
inline void gsl_matrix_free(gsl_matrix* m) { if (m) { delete[] m->data; delete m; } }
inline double gsl_matrix_get(const gsl_matrix* m, size_t i, size_t j) { return m->data[i * m->size2 + j]; }
inline void gsl_matrix_set(gsl_matrix* m, size_t i, size_t j, double x) { m->data[i * m->size2 + j] = x; }

inline double gsl_ran_gaussian_pdf(const double x, const double sigma) {
    const double u = x / std::fabs(sigma);
    const double p = (1 / (std::sqrt(2 * M_PI) * std::fabs(sigma))) * std::exp(-u * u / 2);
    return p;
}

inline unsigned long gsl_rng_get(const gsl_rng* r) { return (unsigned long)const_cast<gsl_rng*>(r)->eng(); }
inline double gsl_rng_uniform(const gsl_rng* r) { return gsl_rng_get(r) / 4294967296.0; }
inline double gsl_rng_uniform_pos(const gsl_rng* r) { double x; do { x = gsl_rng_uniform(r); } while (x == 0); return x; }
inline unsigned long gsl_rng_uniform_int(const gsl_rng* r, unsigned long n) {
    if (n == 0 || n > 0xffffffffUL) gsl_stub_unavailable("gsl_rng_uniform_int with n outside the generator's range");
    const unsigned long scale = 0xffffffffUL / n;
    unsigned long k;
    do { k = gsl_rng_get(r) / scale; } while (k >= n);
    return k;
}
inline double gsl_ran_gaussian(const gsl_rng* r, double sigma) {
    double x, y, r2;
    do { x = -1 + 2 * gsl_rng_uniform_pos(r); y = -1 + 2 * gsl_rng_uniform_pos(r); r2 = x * x + y * y; } while (r2 > 1.0 || r2 == 0);
    return sigma * y * std::sqrt(-2.0 * std::log(r2) / r2);
}
inline gsl_ran_discrete_t* gsl_ran_discrete_preproc(size_t K, const double* P) {
    gsl_ran_discrete_t* g = new gsl_ran_discrete_t; g->K = K; g->cdf = new double[K ? K : 1];
    double acc = 0;
    for (size_t k = 0; k < K; k++) { if (P[k] < 0) gsl_stub_unavailable("gsl_ran_discrete_preproc with a negative weight"); acc += P[k]; g->cdf[k] = acc; }
    return g;
}
inline size_t gsl_ran_discrete(const gsl_rng* r, const gsl_ran_discrete_t* g) {
    const double x = gsl_rng_uniform(r) * g->cdf[g->K - 1];
    size_t lo = 0, hi = g->K - 1;
    while (lo < hi) { const size_t mid = (lo + hi) / 2; if (g->cdf[mid] > x) hi = mid; else lo = mid + 1; }
    return lo;
}
inline void gsl_ran_discrete_free(gsl_ran_discrete_t* g) { if (g) { delete[] g->cdf; delete g; } }
inline int gsl_ran_multivariate_gaussian(const gsl_rng* r, const gsl_vector* mu, const gsl_matrix* L, gsl_vector* result) {
    const size_t n = mu->size;
    double* z = new double[n ? n : 1];
    for (size_t i = 0; i < n; i++) z[i] = gsl_ran_gaussian(r, 1.0);
    for (size_t i = 0; i < n; i++) { double s = 0; for (size_t j = 0; j <= i; j++) s += gsl_matrix_get(L, i, j) * z[j]; result->data[i] = s + mu->data[i]; }
    delete[] z;
    return GSL_SUCCESS;
}
inline int gsl_ran_multivariate_gaussian_vcov(const gsl_matrix* X, gsl_matrix* sigma_hat) {
    const size_t n = X->size1, d = X->size2;
    double* mu = new double[d ? d : 1]();
    for (size_t i = 0; i < n; i++) for (size_t j = 0; j < d; j++) mu[j] += gsl_matrix_get(X, i, j);
    for (size_t j = 0; j < d; j++) mu[j] /= (double)n;
    for (size_t a = 0; a < d; a++) for (size_t b = 0; b < d; b++) {
        double s = 0;
        for (size_t i = 0; i < n; i++) s += (gsl_matrix_get(X, i, a) - mu[a]) * (gsl_matrix_get(X, i, b) - mu[b]);
        gsl_matrix_set(sigma_hat, a, b, s / ((double)n - 1));
    }
    delete[] mu;
    return GSL_SUCCESS;
}
inline int gsl_linalg_cholesky_decomp1(gsl_matrix* A) {
    const size_t n = A->size1;
    for (size_t j = 0; j < n; j++) {
        double d = gsl_matrix_get(A, j, j);
        for (size_t k = 0; k < j; k++) d -= gsl_matrix_get(A, j, k) * gsl_matrix_get(A, j, k);
        if (!(d > 0)) gsl_stub_unavailable("gsl_linalg_cholesky_decomp1 of a matrix that is not positive definite (GSL calls its error handler: abort)");
        const double ljj = std::sqrt(d);
        gsl_matrix_set(A, j, j, ljj);
        for (size_t i = j + 1; i < n; i++) {
            double s = gsl_matrix_get(A, i, j);
            for (size_t k = 0; k < j; k++) s -= gsl_matrix_get(A, i, k) * gsl_matrix_get(A, j, k);
            gsl_matrix_set(A, i, j, s / ljj);
        }
    }
    return GSL_SUCCESS;
}
inline double gsl_sf_lnchoose(unsigned int, unsigned int) { gsl_stub_unavailable("gsl_sf_lnchoose"); }
static const gsl_multimin_fminimizer_type gsl_stub_nmsimplex2_type = {"nmsimplex2 (stand-in)"};
static const gsl_multimin_fminimizer_type* const gsl_multimin_fminimizer_nmsimplex2 = &gsl_stub_nmsimplex2_type;
inline gsl_multimin_fminimizer* gsl_multimin_fminimizer_alloc(const gsl_multimin_fminimizer_type*, size_t) { gsl_stub_unavailable("gsl_multimin_fminimizer_alloc"); }
inline int gsl_multimin_fminimizer_set(gsl_multimin_fminimizer*, gsl_multimin_function*, const gsl_vector*, const gsl_vector*) { gsl_stub_unavailable("gsl_multimin_fminimizer_set"); }
inline int gsl_multimin_fminimizer_iterate(gsl_multimin_fminimizer*) { gsl_stub_unavailable("gsl_multimin_fminimizer_iterate"); }
inline double gsl_multimin_fminimizer_size(const gsl_multimin_fminimizer*) { gsl_stub_unavailable("gsl_multimin_fminimizer_size"); }
inline int gsl_multimin_test_size(double, double) { gsl_stub_unavailable("gsl_multimin_test_size"); }
inline void gsl_multimin_fminimizer_free(gsl_multimin_fminimizer*) { gsl_stub_unavailable("gsl_multimin_fminimizer_free"); }
}
