// sharded_test.cpp — a C++ host (no Python, no torch) drives the row-sharded weight update through the C ABI the way a C++ AbcSmc
// would (src/AbcSmc.cpp:1053-1064): abcb200_group_create over n GPUs, abcb200_weights_sharded on host buffers, compared with
// abcb200_weights on one GPU. Usage: sharded_test n_gpus N_new N_old P [out.bin]   (out.bin: inputs + both results for the checker)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/abcsmc_b200.h"

static uint64_t sm64(uint64_t& s) { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
static double unif(uint64_t& s) { return (double)(sm64(s) >> 11) * (1.0 / 9007199254740992.0); }
static double gauss(uint64_t& s) { const double u1 = unif(s) + 1e-300, u2 = unif(s); return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2); }

int main(int argc, char** argv) {
    const int n_gpus = argc > 1 ? atoi(argv[1]) : 0;
    const int64_t N_new = argc > 2 ? atoll(argv[2]) : 3001, N_old = argc > 3 ? atoll(argv[3]) : 2000;
    const int P = argc > 4 ? atoi(argv[4]) : 30;
    uint64_t seed = 20261018;
    std::vector<double> th_old((size_t)N_old * P), th_new((size_t)N_new * P), w_old((size_t)N_old), dv((size_t)P), numer((size_t)N_new);
    for (int p = 0; p < P; p++) {
        double m = 0, m2 = 0;
        for (int64_t j = 0; j < N_old; j++) { const double v = 0.5 + 0.1 * gauss(seed); th_old[(size_t)p * N_old + j] = v; m += v; }
        m /= (double)N_old;
        for (int64_t j = 0; j < N_old; j++) { const double d = th_old[(size_t)p * N_old + j] - m; m2 += d * d; }
        dv[(size_t)p] = 2.0 * m2 / (double)(N_old - 1);
    }
    double ss = 0;
    for (int64_t j = 0; j < N_old; j++) { w_old[(size_t)j] = 0.5 + unif(seed); ss += w_old[(size_t)j] * w_old[(size_t)j]; }
    for (int64_t j = 0; j < N_old; j++) w_old[(size_t)j] /= std::sqrt(ss);
    for (int64_t i = 0; i < N_new; i++) {
        const int64_t pick = (int64_t)(unif(seed) * (double)N_old);
        for (int p = 0; p < P; p++) th_new[(size_t)p * N_new + i] = th_old[(size_t)p * N_old + pick] + std::sqrt(dv[(size_t)p]) * gauss(seed);
        numer[(size_t)i] = 0.25 + unif(seed);
    }
    abcb200_group* g = nullptr;
    int rc = abcb200_group_create(n_gpus, nullptr, &g);
    if (rc != ABCB200_OK) { fprintf(stderr, "abcb200_group_create(%d) failed: %d\n", n_gpus, rc); return 2; }
    const int G = abcb200_group_size(g);
    std::vector<double> w_sh((size_t)N_new), w_one((size_t)N_new);
    double worst = 0;
    for (int algo = 0; algo <= 2; algo++) {
        rc = abcb200_weights_sharded(g, numer.data(), th_new.data(), N_new, N_new, th_old.data(), N_old, N_old, w_old.data(), dv.data(), P, algo, w_sh.data());
        if (rc != ABCB200_OK) { fprintf(stderr, "abcb200_weights_sharded: %d %s\n", rc, abcb200_group_last_error(g)); return 3; }
        abcb200_ctx* c0 = abcb200_group_ctx(g, 0);
        rc = abcb200_weights(c0, numer.data(), th_new.data(), N_new, N_new, th_old.data(), N_old, N_old, w_old.data(), dv.data(), P, algo, w_one.data());
        if (rc != ABCB200_OK) { fprintf(stderr, "abcb200_weights: %d %s\n", rc, abcb200_last_error(c0)); return 3; }
        double e = 0, nrm = 0;
        for (int64_t i = 0; i < N_new; i++) { e = std::fmax(e, std::fabs(w_sh[(size_t)i] - w_one[(size_t)i]) / std::fabs(w_one[(size_t)i])); nrm += w_sh[(size_t)i] * w_sh[(size_t)i]; }
        printf("algo %d: %d GPU(s), max rel diff sharded vs single = %.3e, |w|^2 = %.15f\n", algo, G, e, nrm);
        worst = std::fmax(worst, e);
        if (!(std::fabs(nrm - 1.0) < 1e-12)) { fprintf(stderr, "not L2-normalised\n"); return 4; }
    }
    if (argc > 5) {
        FILE* o = fopen(argv[5], "wb");
        if (!o) { perror("out"); return 2; }
        fwrite(th_new.data(), 8, th_new.size(), o); fwrite(th_old.data(), 8, th_old.size(), o); fwrite(w_old.data(), 8, w_old.size(), o);
        fwrite(dv.data(), 8, dv.size(), o); fwrite(numer.data(), 8, numer.size(), o); fwrite(w_sh.data(), 8, w_sh.size(), o);
        fclose(o);
    }
    abcb200_group_destroy(g);
    if (!(worst < 1e-11)) { fprintf(stderr, "sharded result differs from the single-GPU result: %.3e\n", worst); return 5; }
    printf("sharded ok: G=%d N_new=%lld N_old=%lld P=%d\n", G, (long long)N_new, (long long)N_old, P);
    return 0;
}
