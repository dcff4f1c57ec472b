// TEST INFRASTRUCTURE (oracle/): declarations of the GSL entry points the reference's src/AbcUtil.cpp and headers name, so that
// the UNMODIFIED reference source compiles in this image (GSL is absent, there is no network). Implemented for real, from GSL's
// documented behaviour: gsl_vector / gsl_matrix storage and gsl_ran_gaussian_pdf (randist/gauss.c: u = x / fabs(sigma);
// p = (1 / (sqrt(2 pi) fabs(sigma))) exp(-u u / 2)) — the only GSL arithmetic on the hot path (AbcUtil.cpp:574, Priors.h:54).
// Everything that needs GSL's random streams, special functions or minimisers aborts with a message when called: those are
// outside the path this repository pins (proposal sampling has distributional parity only).
#pragma once
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cmath>

#define GSL_SUCCESS 0
#define GSL_CONTINUE (-2)

extern "C++" {
struct gsl_rng { unsigned long long state; };
struct gsl_vector { size_t size; double* data; };
struct gsl_matrix { size_t size1, size2; double* data; };     // row-major like GSL
struct gsl_ran_discrete_t { size_t K; };
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

inline double gsl_ran_gaussian(const gsl_rng*, double) { gsl_stub_unavailable("gsl_ran_gaussian"); }
inline double gsl_rng_uniform(const gsl_rng*) { gsl_stub_unavailable("gsl_rng_uniform"); }
inline unsigned long gsl_rng_uniform_int(const gsl_rng*, unsigned long) { gsl_stub_unavailable("gsl_rng_uniform_int"); }
inline gsl_ran_discrete_t* gsl_ran_discrete_preproc(size_t, const double*) { gsl_stub_unavailable("gsl_ran_discrete_preproc"); }
inline size_t gsl_ran_discrete(const gsl_rng*, const gsl_ran_discrete_t*) { gsl_stub_unavailable("gsl_ran_discrete"); }
inline void gsl_ran_discrete_free(gsl_ran_discrete_t*) { gsl_stub_unavailable("gsl_ran_discrete_free"); }
inline int gsl_ran_multivariate_gaussian(const gsl_rng*, const gsl_vector*, const gsl_matrix*, gsl_vector*) { gsl_stub_unavailable("gsl_ran_multivariate_gaussian"); }
inline int gsl_ran_multivariate_gaussian_vcov(const gsl_matrix*, gsl_matrix*) { gsl_stub_unavailable("gsl_ran_multivariate_gaussian_vcov"); }
inline int gsl_linalg_cholesky_decomp1(gsl_matrix*) { gsl_stub_unavailable("gsl_linalg_cholesky_decomp1"); }
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
