// dropin_refhdr_test.cpp — INTEGRATION.md §2 against the reference's REAL headers: <AbcSmc/AbcUtil.h>, <AbcSmc/Priors.h> and (through
// them) <PLS/pls.h> are included from /root/reference unmodified (Eigen / GSL come from the stand-ins of oracle/shim/, the only way
// they compile in this image), then abc_b200.hpp with ABCB200_DROP_IN supplies the bodies of the five ABC:: functions those headers
// DECLARE (include/AbcSmc/AbcUtil.h:103, 146-172). The call sites below use the reference's declarations, types (Mat2D, Row, Col,
// float_type) and prior classes (ABC::ContinuousUniformPrior, Priors.h:86-110), so the program only links if the drop-in's
// signatures are exactly the reference's. src/AbcUtil.cpp is NOT linked: every ABC:: call lands in the CUDA library.
// Reads the binary case of tests/test_cpp_adapter.py, writes results for the checker. Built only where /root/reference exists;
// the executable (tests/cpp/_build/, git-ignored) travels to the GPU box.
#include <cstdio>
#include <cstdlib>

#include <AbcSmc/AbcUtil.h>
#include <AbcSmc/Priors.h>
#define ABCB200_DROP_IN
#include "../../abcsmc_b200/host/abc_b200.hpp"

static void rd(FILE* f, void* p, size_t n) { if (fread(p, 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }
template <class T> static void wr(FILE* o, const T* p, size_t n) { fwrite(p, sizeof(T), n, o); }

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: dropin_refhdr_test case.bin out.bin\n"); return 2; }
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

    // the statements of AbcSmc::read_SMC_sets_from_database / calculate_predictive_prior_weights, with the reference's own types
    std::vector<size_t> order = ABC::particle_ranking_PLS(met, par, target, 0.5);               // AbcSmc.cpp:635-637
    order.resize((size_t)Npp);                                                                   // AbcSmc.cpp:645-646
    const Mat2D post = par(order, Eigen::placeholders::all);                                     // AbcSmc.cpp:648
    const Row dv = ABC::calculate_doubled_variance(post);                                        // AbcSmc.cpp:1045
    std::vector<ABC::ContinuousUniformPrior> pars((size_t)P, ABC::ContinuousUniformPrior("theta", "th", 0.0, 2.0));
    std::vector<const ABC::Parameter*> mpars;
    for (auto& p : pars) mpars.push_back(&p);
    const Row w0 = ABC::weight_predictive_prior(mpars, post);                                    // AbcSmc.cpp:1050
    const Row w = ABC::weight_predictive_prior(mpars, post, th_old, w_old, dv_old);              // AbcSmc.cpp:1056-1063
    std::vector<size_t> simple = ABC::particle_ranking_simple(met, par, target);                 // AbcSmc.cpp:638-640
    simple.resize((size_t)Npp);
    const Col dist = ABC::euclidean(post, dv);                                                   // any (N_pp x P, P) pair

    FILE* o = fopen(argv[2], "wb");
    if (!o) { perror("out"); return 2; }
    std::vector<long> ord(order.begin(), order.end()), simp(simple.begin(), simple.end());
    wr(o, ord.data(), ord.size()); wr(o, dv.data(), (size_t)P); wr(o, w0.data(), (size_t)Npp); wr(o, w.data(), (size_t)Npp);
    wr(o, simp.data(), simp.size()); wr(o, dist.data(), (size_t)Npp);
    fclose(o);
    printf("dropin_refhdr ok: N=%ld K=%ld P=%ld\n", N, K, P);
    return 0;
}
