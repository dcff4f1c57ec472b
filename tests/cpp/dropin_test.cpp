// dropin_test.cpp — compiles abc_b200.hpp with ABCB200_DROP_IN + ABCB200_DROP_IN_SAMPLER + ABCB200_DROP_IN_PLS against stand-ins of
// the reference's typedefs (ref_stub.hpp) and drives (a) namespace ABC the way AbcSmc.cpp:634-664, 1041-1066 does and (b) namespace
// PLS the way lib/PLS/src/main.cpp:19-41 does. Reads the binary case of tests/test_cpp_adapter.py, writes results for the checker.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <sstream>

#include "ref_stub.hpp"
#define ABCB200_DROP_IN
#define ABCB200_DROP_IN_SAMPLER
#define ABCB200_DROP_IN_PLS
#include "../../abcsmc_b200/host/abc_b200.hpp"

static void rd(FILE* f, void* p, size_t n) { if (fread(p, 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }
template <class T> static void wr(FILE* o, const T* p, size_t n) { fwrite(p, sizeof(T), n, o); }

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: dropin_test case.bin out.bin\n"); return 2; }
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

    // ---- (a) namespace ABC, as AbcSmc::read_SMC_sets_from_database calls it ---------------------------------------------------
    std::vector<size_t> order = ABC::particle_ranking_PLS(met, par, target, 0.5);               // AbcSmc.cpp:635-637
    order.resize((size_t)Npp);                                                                   // AbcSmc.cpp:645-646
    Mat2D post(Npp, P);                                                                          // AbcSmc.cpp:648
    for (long j = 0; j < P; j++) for (long i = 0; i < Npp; i++) post(i, j) = par((long)order[(size_t)i], j);
    const Row dv = ABC::calculate_doubled_variance(post);                                        // AbcSmc.cpp:1045
    std::vector<ContinuousUniformPrior> pars((size_t)P, ContinuousUniformPrior(0.0, 2.0));
    std::vector<const Parameter*> mpars;
    for (auto& p : pars) mpars.push_back(&p);
    const Row w0 = ABC::weight_predictive_prior(mpars, post);                                    // AbcSmc.cpp:1050
    const Row w = ABC::weight_predictive_prior(mpars, post, th_old, w_old, dv_old);              // AbcSmc.cpp:1056-1063
    const std::vector<size_t> simple_full = ABC::particle_ranking_simple(met, par, target);
    const Col dist = ABC::euclidean(post, dv);                                                    // any (N_pp x P, P) pair
    gsl_rng rng{12345};
    Col wc(Npp);
    for (long i = 0; i < Npp; i++) wc[i] = w[i];
    const Mat2D prop = ABC::sample_predictive_priors(&rng, (size_t)Npp, wc, post, mpars, dv);    // AbcSmc.cpp:508-515

    // the same set through the chained entry point (set 0: weights 1 / n; then the set again as "set 1" against itself)
    {
        ABC_B200::SmcChain<Mat2D, Row> chain((int)P);
        auto r0 = chain.process_set(met, par, target, mpars, true, 0.5, (size_t)Npp);
        auto r1 = chain.process_set(met, par, target, mpars, true, 0.5, (size_t)Npp);
        if (chain.sets() != 2) { fprintf(stderr, "chain sets\n"); return 3; }
        for (long i = 0; i < Npp; i++) if (r0.predictive_prior[(size_t)i] != order[(size_t)i] || r1.predictive_prior[(size_t)i] != order[(size_t)i] || r0.weights[i] != w0[i]) { fprintf(stderr, "chain order / set-0 weights\n"); return 3; }
        for (long j = 0; j < P; j++) if (r0.doubled_variance[j] != dv[j]) { fprintf(stderr, "chain dv\n"); return 3; }
        const Row w_self = ABC::weight_predictive_prior(mpars, post, post, w0, dv);             // set 1 against set 0 = itself
        for (long i = 0; i < Npp; i++) if (std::fabs(r1.weights[i] - w_self[i]) > 1e-12 * std::fabs(w_self[i])) { fprintf(stderr, "chain weights %ld %.17g %.17g\n", i, r1.weights[i], w_self[i]); return 3; }
    }

    // ---- (b) namespace PLS, as lib/PLS/src/main.cpp does ---------------------------------------------------------------------------
    using namespace PLS;
    const long Nl = N < 240 ? N : 240, Nh = N - Nl < 500 ? N - Nl : 500;
    Mat2D Xo(Nl, K), Yo(Nl, P), Xn(Nh, K), Yn(Nh, P);
    for (long j = 0; j < K; j++) { for (long i = 0; i < Nl; i++) Xo(i, j) = met(i, j); for (long i = 0; i < Nh; i++) Xn(i, j) = met(Nl + i, j); }
    for (long j = 0; j < P; j++) { for (long i = 0; i < Nl; i++) Yo(i, j) = par(i, j); for (long i = 0; i < Nh; i++) Yn(i, j) = par(Nl + i, j); }
    const Mat2D X = colwise_z_scores(Xo), Y = colwise_z_scores(Yo);
    const size_t ncomp = (size_t)(K < 6 ? K : 6);
    Model plsm(X, Y, METHOD::KERNEL_TYPE1, ncomp);
    Model copy_of(plsm);                                        // value semantics, as the reference's struct
    std::ostringstream os;
    copy_of.print_state(os);
    plsm.print_explained_variance(X, Y, os);
    Residual looerror = plsm.cv_LOO();
    print_validation(looerror, MSE, os);
    std::mt19937 gen(777);
    Residual lsoerror = plsm.cv_LSO(0.3, 4, gen);
    print_validation(lsoerror, MSE, os);
    Residual nderror = plsm.cv_NEW_DATA(Xn, Yn);
    const Mat2D press_nd = validation(nderror, RESS);
    const Colsz nc_nd = optimal_num_components(nderror);           // streamed (ALPHA = 0.1)
    const Colsz nc_nd05 = optimal_num_components(nderror, 0.05);   // through the materialised cube
    const std::vector<Mat2D> ev_loo = looerror.errors(), ev_nd = nderror.errors(), ev_lso = lsoerror.errors();
    const Mat2D mse_loo = validation(looerror, MSE);
    const Colsz nc_loo = optimal_num_components(looerror);
    const Row ev = plsm.explained_variance(X, Y);
    if (looerror.method() != "LOO" || lsoerror.method() != "LSO" || nderror.method() != "NEW DATA") { fprintf(stderr, "labels\n"); return 3; }
    if (os.str().find("coefficients:") == std::string::npos || os.str().find("LOO Validation:") == std::string::npos) { fprintf(stderr, "print\n"); return 3; }

    FILE* o = fopen(argv[2], "wb");
    if (!o) { perror("out"); return 2; }
    std::vector<long> ord(order.begin(), order.end()), smp(simple_full.begin(), simple_full.begin() + Npp);
    wr(o, ord.data(), ord.size()); wr(o, dv.data(), (size_t)P); wr(o, w0.data(), (size_t)Npp); wr(o, w.data(), (size_t)Npp); wr(o, smp.data(), smp.size());
    wr(o, dist.data(), (size_t)Npp); wr(o, prop.data(), (size_t)Npp * P);
    const long shp[4] = {Nl, Nh, (long)ncomp, (long)(ev_lso[0].rows())};
    wr(o, shp, 4);
    for (long y = 0; y < P; y++) wr(o, ev_loo[(size_t)y].data(), (size_t)Nl * ncomp);
    wr(o, mse_loo.data(), (size_t)P * ncomp); wr(o, nc_loo.data(), (size_t)P);
    for (long y = 0; y < P; y++) wr(o, ev_nd[(size_t)y].data(), (size_t)Nh * ncomp);
    wr(o, press_nd.data(), (size_t)P * ncomp); wr(o, nc_nd.data(), (size_t)P); wr(o, nc_nd05.data(), (size_t)P);
    wr(o, ev.data(), (size_t)P);
    fclose(o);
    printf("dropin ok: N=%ld K=%ld P=%ld top=%ld\n", N, K, P, Npp);
    return 0;
}
