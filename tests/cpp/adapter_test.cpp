// adapter_test.cpp — drives abcsmc_b200/host/abc_b200.hpp with a minimal Eigen-like column-major matrix type, the way
// AbcSmc.cpp:634-664 and :1041-1066 drive namespace ABC: rank -> truncate -> gather -> doubled variance -> weights.
// Reads a binary case written by tests/test_cpp_adapter.py and writes the results next to it; the test compares them with
// the CPU oracle. Build: g++ -std=c++17 -I. tests/cpp/adapter_test.cpp -Labcsmc_b200 -labcsmc_b200 -Wl,-rpath,...
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <random>
#include <vector>

#include "../../abcsmc_b200/host/abc_b200.hpp"

struct Mat2D {   // what the adapter needs from Eigen::MatrixXd
    Mat2D(long r, long c) : r_(r), c_(c), d_((size_t)r * c) {}
    double* data() { return d_.data(); }
    const double* data() const { return d_.data(); }
    long rows() const { return r_; }
    long cols() const { return c_; }
    long outerStride() const { return r_; }
    double& operator()(long i, long j) { return d_[(size_t)j * r_ + i]; }
    double operator()(long i, long j) const { return d_[(size_t)j * r_ + i]; }
    long r_, c_;
    std::vector<double> d_;
};
struct Row {     // Eigen::RowVectorXd / VectorXd
    explicit Row(long n) : d_((size_t)n) {}
    double* data() { return d_.data(); }
    const double* data() const { return d_.data(); }
    long size() const { return (long)d_.size(); }
    double& operator[](long i) { return d_[(size_t)i]; }
    double operator[](long i) const { return d_[(size_t)i]; }
    std::vector<double> d_;
};
struct Parameter {   // include/AbcSmc/Parameter.h:58 + Priors.h:86-110 — a ContinuousUniformPrior on [a, b]
    Parameter(double a, double b) : a_(a), b_(b) {}
    virtual ~Parameter() {}
    virtual double likelihood(double v) const { return (v >= a_ && v <= b_) ? 1.0 / (b_ - a_) : 0.0; }   // Priors.h:101-103
    virtual double recast(double v) const { return v; }                                                   // Priors.h:105
    bool valid(double v) const { return likelihood(v) != 0.0; }                                           // Parameter.h:77
    virtual double get_mean() const { return (a_ + b_) / 2.0; }                                           // Priors.h:92
    virtual double get_sd() const { return (b_ - a_) / std::sqrt(12.0); }                                 // Priors.h:93
    double a_, b_;
};

struct abcb200_adapter_flat_check { double lo, hi; };
static abcb200_adapter_flat_check check_flatten(const Parameter& p) {      // the bounds recovered from mean / sd are the exact ones
    const ABC_B200::FlatPrior f = ABC_B200::flatten_prior(p);
    if (f.lo != p.a_ || f.hi != p.b_ || f.integral != 0 || f.mean != (p.a_ + p.b_) / 2.0) { fprintf(stderr, "flatten_prior: [%g, %g] int=%d\n", f.lo, f.hi, f.integral); exit(3); }
    return {f.lo, f.hi};
}
static void rd(FILE* f, void* p, size_t n) { if (fread(p, 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: adapter_test case.bin out.bin\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("case"); return 2; }
    long hdr[5];   // N, K, P, N_pp, N_old
    rd(f, hdr, sizeof(hdr));
    const long N = hdr[0], K = hdr[1], P = hdr[2], Npp = hdr[3], Nold = hdr[4];
    Mat2D met(N, K), par(N, P), th_old(Nold, P);
    Row target(K), w_old(Nold), dv_old(P);
    rd(f, met.data(), sizeof(double) * N * K); rd(f, par.data(), sizeof(double) * N * P); rd(f, target.data(), sizeof(double) * K);
    rd(f, th_old.data(), sizeof(double) * Nold * P); rd(f, w_old.data(), sizeof(double) * Nold); rd(f, dv_old.data(), sizeof(double) * P);
    fclose(f);

    std::vector<size_t> order = ABC_B200::particle_ranking_PLS(met, par, target, 0.5);        // AbcSmc.cpp:635-637
    order.resize((size_t)Npp);                                                                  // AbcSmc.cpp:645-646
    Mat2D post(Npp, P);                                                                         // AbcSmc.cpp:648 (fancy indexing)
    for (long j = 0; j < P; j++) for (long i = 0; i < Npp; i++) post(i, j) = par((long)order[(size_t)i], j);
    const Row dv = ABC_B200::calculate_doubled_variance<Row>(post);                            // AbcSmc.cpp:1045
    std::vector<Parameter> pars((size_t)P, Parameter(0.0, 2.0));
    std::vector<const Parameter*> mpars;
    for (auto& p : pars) mpars.push_back(&p);
    const Row w0 = ABC_B200::weight_predictive_prior<Row>(mpars, post);                        // AbcSmc.cpp:1050 (set 0)
    const Row w = ABC_B200::weight_predictive_prior<Row>(mpars, post, th_old, w_old, dv_old);  // AbcSmc.cpp:1056-1063
    const std::vector<size_t> simple = ABC_B200::particle_ranking_simple(met, par, target, (size_t)Npp);
    PLS_B200::Model<Mat2D, Row> model(met, par, PLS_B200::KERNEL_TYPE1, (size_t)K);             // un-standardised on purpose: any X, Y
    const Mat2D B = model.coefficients();
    // Model::cv_LOO / cv_LSO through the wrapper, on the first 300 rows (the oracle refits N times)
    const long Nl = N < 300 ? N : 300;
    Mat2D metl(Nl, K), parl(Nl, P);
    for (long j = 0; j < K; j++) for (long i = 0; i < Nl; i++) metl(i, j) = met(i, j);
    for (long j = 0; j < P; j++) for (long i = 0; i < Nl; i++) parl(i, j) = par(i, j);
    PLS_B200::Model<Mat2D, Row> modl(metl, parl, PLS_B200::KERNEL_TYPE1, (size_t)K);
    const PLS_B200::Validation<Mat2D> loo = modl.cv_LOO(metl, parl);
    std::mt19937 rng(12345);
    const PLS_B200::Validation<Mat2D> lso = modl.cv_LSO(metl, parl, 0.2, 3, rng);
    std::vector<long> shuf((size_t)(3 * Nl));                 // the same partitions, for the checker (rand_nchoosek, pls.cpp:217-227)
    {
        std::mt19937 rng2(12345);
        std::vector<long> full((size_t)Nl);
        for (long i = 0; i < Nl; i++) full[(size_t)i] = i;
        for (int t = 0; t < 3; t++) { std::shuffle(full.begin(), full.end(), rng2); std::copy(full.begin(), full.end(), shuf.begin() + (std::ptrdiff_t)t * Nl); }
    }
    // next-set proposals from the predictive prior just built (AbcSmc.cpp:508-515): 2 * Npp samples, uniform priors on [0, 2]
    const abcb200_adapter_flat_check fc = check_flatten(pars[0]);
    const Mat2D prop = ABC_B200::sample_predictive_priors<Mat2D>(20240517ull, (size_t)(2 * Npp), w, post, mpars, dv);

    FILE* o = fopen(argv[2], "wb");
    if (!o) { perror("out"); return 2; }
    std::vector<long> ord(order.begin(), order.end()), smp(simple.begin(), simple.end());
    fwrite(ord.data(), sizeof(long), ord.size(), o);
    fwrite(dv.data(), sizeof(double), (size_t)P, o);
    fwrite(w0.data(), sizeof(double), (size_t)Npp, o);
    fwrite(w.data(), sizeof(double), (size_t)Npp, o);
    fwrite(smp.data(), sizeof(long), smp.size(), o);
    fwrite(B.data(), sizeof(double), (size_t)K * P, o);
    fwrite(prop.data(), sizeof(double), (size_t)(2 * Npp) * P, o);
    std::vector<long> nloo(loo.num_components.begin(), loo.num_components.end()), nlso(lso.num_components.begin(), lso.num_components.end());
    fwrite(nloo.data(), sizeof(long), nloo.size(), o);
    fwrite(loo.press.data(), sizeof(double), (size_t)P * K, o);
    fwrite(nlso.data(), sizeof(long), nlso.size(), o);
    fwrite(lso.press.data(), sizeof(double), (size_t)P * K, o);
    fwrite(shuf.data(), sizeof(long), shuf.size(), o);
    (void)fc;
    fclose(o);
    printf("adapter ok: N=%ld K=%ld P=%ld top=%ld\n", N, K, P, Npp);
    return 0;
}
