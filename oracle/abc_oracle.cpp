// abc_oracle.cpp — CPU restatement of AbcSmc's per-set post-simulation hot path.
//
// TEST INFRASTRUCTURE ONLY. This file is the parity checker for the CUDA path in
// abcsmc_b200/csrc/. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs may load it. The product (libabcsmc_b200.so) never links or calls it.
//
// PARITY STATUS: pinned against the reference's own source files compiled here with stand-in headers for the two absent
// third-party libraries; NOT pinned against a build with real Eigen / GSL ("parity unpinned" in that strict sense: the
// reference (tjhladish/AbcSmc @ 2ea44e7 with PLS submodule @ d976786) cannot be built as shipped in this image — Eigen >= 3.4.90
// (un-vendored submodule lib/PLS/lib/eigen @ 23e1541) and GSL are absent, and there is no network).
// This file restates the reference's algorithm, function by function, in dependency-free C++17, following the reference's
// loop and operation order where that is cheap. It is pinned by
//   (o)  oracle/_ref/libabcref.so (oracle/Makefile `ref`): the reference's UNMODIFIED lib/PLS/src/pls.cpp and src/AbcUtil.cpp,
//        compiled where they lie under /root/reference against oracle/shim/ (an eager stand-in for the subset of Eigen's API those
//        files use, with a tred2/tql2 eigen-solver; gsl_ran_gaussian_pdf from GSL's formula). tests/test_ref_pin.py compares this
//        file with it function by function (orders and component counts bit-exact, FP64 <= 1e-10; observed <= 1e-13), and
//        tests/golden/ref_small.npz / ref_fullsize_C3.npz hold its outputs (make_ref_fixtures.py) — at the full dengue shape
//        (N=250k, K=150, P=30) the first 5000 ranks from the reference's code equal this file's;
//   (i)  the reference's own three known-answer tests (tests/abcutil.cpp:11-40, tests/pls.cpp:15-24),
//   (ii) an independent numpy/scipy formulation (tests/np_reference.py) on the reference's toy
//        fixtures (lib/PLS/toyX.csv, toyY.csv, nir.csv, octane.csv) and seeded synthetic data,
//   (iii) algebraic invariants (OLS limit, score orthogonality, PRESS prefix identity, weight scale
//        invariance) — see tests/test_oracle.py.
//
// Third-party arithmetic restated here (absent from /root/reference):
//   * Eigen::EigenSolver<MatrixXd> at lib/PLS/src/pls.cpp:406 — applied by the reference to the
//     symmetric PSD matrix XY^T XY. Restated as a cyclic Jacobi eigen-solver (symmetric input, so
//     eigenpairs are real and agree with any backward-stable general solver up to sign and O(eps/gap)).
//   * gsl_ran_gaussian_pdf(x, sigma) at src/AbcUtil.cpp:574 and include/AbcSmc/Priors.h:54 —
//     GSL randist/gauss.c: u = x/fabs(sigma); p = (1/(sqrt(2*pi)*fabs(sigma)))*exp(-u*u/2).
//   * std::sort (libstdc++ introsort) at lib/PLS/include/PLS/pls.h:62 — used as is, same comparator.
//
// All matrices are column-major (Eigen::MatrixXd default, lib/PLS/include/PLS/pls.h:22-27) with
// leading dimension == rows. Build: g++ -O2 -std=c++17 -fPIC -shared (the reference's flags,
// CMakeLists.txt:5-6; no -march, so no FMA contraction on x86-64).

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

namespace orc {

typedef double float_type;

struct Mat {
    long r = 0, c = 0;
    std::vector<double> d;
    Mat() {}
    Mat(long rows, long cols) : r(rows), c(cols), d((size_t)rows * cols, 0.0) {}
    Mat(long rows, long cols, const double* src, long ld) : r(rows), c(cols), d((size_t)rows * cols) {
        for (long j = 0; j < cols; j++) std::memcpy(&d[(size_t)j * rows], src + (size_t)j * ld, sizeof(double) * rows);
    }
    double& operator()(long i, long j) { return d[(size_t)j * r + i]; }
    const double& operator()(long i, long j) const { return d[(size_t)j * r + i]; }
    double* col(long j) { return &d[(size_t)j * r]; }
    const double* col(long j) const { return &d[(size_t)j * r]; }
    Mat rows_range(long r0, long n) const {   // Eigen topRows/bottomRows materialised
        Mat out(n, c);
        for (long j = 0; j < c; j++) std::memcpy(out.col(j), col(j) + r0, sizeof(double) * n);
        return out;
    }
};
typedef std::vector<double> Vec;

// ---------------------------------------------------------------------------------------------
// lib/PLS/include/PLS/pls.h:58-69 — ordered(): index sort, strict <, std::sort (not stable)
static std::vector<size_t> ordered(const double* v, size_t n) {
    std::vector<size_t> result(n);
    std::iota(result.begin(), result.end(), 0);
    std::sort(result.begin(), result.end(), [v](const size_t& lhs, const size_t& rhs) { return v[lhs] < v[rhs]; });
    return result;
}

// colwise().mean() as used at src/AbcUtil.cpp:432 and lib/PLS/src/pls.cpp:108
static Vec colwise_mean(const Mat& m) {
    Vec mu(m.c, 0.0);
    for (long j = 0; j < m.c; j++) {
        const double* x = m.col(j);
        double s = 0;
        for (long i = 0; i < m.r; i++) s += x[i];
        mu[j] = s / (double)m.r;
    }
    return mu;
}

// lib/PLS/src/pls.cpp:69-73 — SST(mat, means): N<2 -> zeros; sum((x-mean)^2) by column
static Vec SST(const Mat& m, const Vec& means) {
    Vec out(m.c, 0.0);
    if (m.r < 2) return out;
    for (long j = 0; j < m.c; j++) {
        const double* x = m.col(j);
        double s = 0;
        for (long i = 0; i < m.r; i++) { const double dlt = x[i] - means[j]; s += dlt * dlt; }
        out[j] = s;
    }
    return out;
}

// lib/PLS/src/pls.cpp:79-83 — colwise_stdev: sqrt(SST/(N-1))
static Vec colwise_stdev(const Mat& m, const Vec& means) {
    const double N = (double)m.r;
    Vec s = SST(m, means);
    for (auto& v : s) v = std::sqrt(v / (N - 1));
    return s;
}

// lib/PLS/src/pls.cpp:89-91 — z_scores(obs, mean, stdev): no zero guard
static Vec z_scores(const Vec& obs, const Vec& mean, const Vec& sd) {
    Vec z(obs.size());
    for (size_t j = 0; j < obs.size(); j++) z[j] = (obs[j] - mean[j]) / sd[j];
    return z;
}

// lib/PLS/src/pls.cpp:93-105 — colwise_z_scores(mat, mean, stdev).
// The reference computes a zero-guarded local_sd (:94-100) but divides by the UNGUARDED stdev (:103),
// so constant columns give 0/0 = NaN. Restated literally.
static Mat colwise_z_scores(const Mat& m, const Vec& mean, const Vec& sd) {
    Mat z(m.r, m.c);
    for (long j = 0; j < m.c; j++) {
        const double* x = m.col(j);
        double* o = z.col(j);
        for (long i = 0; i < m.r; i++) o[i] = (x[i] - mean[j]) / sd[j];
    }
    return z;
}

// lib/PLS/src/pls.cpp:152-160 — normalcdf: 4-term Abramowitz-Stegun approximation (not erf)
static double normalcdf(const double z) {
    const double c1 = 0.196854, c2 = 0.115194, c3 = 0.000344, c4 = 0.019527;
    const double zstar = std::fabs(z);
    double p = 0.5 / std::pow(1 + c1 * zstar + c2 * zstar * zstar + c3 * zstar * zstar * zstar + c4 * zstar * zstar * zstar * zstar, 4);
    return z < 0 ? p : 1.0 - p;
}

// lib/PLS/src/pls.cpp:190-211 — wilcoxon signed-rank p-value
static double wilcoxon(const double* err_1, const double* err_2, size_t n) {
    Vec del(n), adel(n);
    std::vector<int> sdel(n, 0);
    for (size_t i = 0; i < n; i++) {
        del[i] = std::fabs(err_1[i]) - std::fabs(err_2[i]);
        sdel[i] = (0 < del[i]) - (del[i] < 0);
        adel[i] = std::fabs(del[i]);
    }
    auto s = ordered(adel.data(), n);
    double d = 0;
    for (size_t i = 0; i < n; i++) d += (double)(i + 1) * sdel[s[i]];
    double t = (double)(n * (n + 1)) / 2.0;
    double v = (t - d) / 2.0;
    double ev = t / 2.0;
    double sv = std::sqrt((double)(n * (n + 1) * (2 * n + 1)) / 24.0);   // size_t product, pls.cpp:206
    double z = (v - ev) / sv;
    return 1.0 - normalcdf(z);
}

// Eigen::EigenSolver stand-in for a symmetric matrix (see header): cyclic Jacobi, returns the unit
// eigenvector of the eigenvalue with largest |.| (first on ties; lib/PLS/src/pls.cpp:113-141).
// Sign convention (Eigen's is arbitrary): component of largest magnitude made positive.
static Vec dominant_eigenvector_sym(const Mat& Sin) {
    const long n = Sin.r;
    Mat A = Sin, V(n, n);
    for (long i = 0; i < n; i++) V(i, i) = 1.0;
    for (int sweep = 0; sweep < 100; sweep++) {
        double off = 0, diag = 0;
        for (long p = 0; p < n; p++) { diag += A(p, p) * A(p, p); for (long q = p + 1; q < n; q++) off += A(p, q) * A(p, q); }
        if (off == 0.0 || off <= 1e-34 * diag) break;
        for (long p = 0; p < n - 1; p++) for (long q = p + 1; q < n; q++) {
            const double apq = A(p, q);
            if (apq == 0.0) continue;
            const double theta = (A(q, q) - A(p, p)) / (2.0 * apq);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
            const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
            for (long k = 0; k < n; k++) { const double akp = A(k, p), akq = A(k, q); A(k, p) = c * akp - s * akq; A(k, q) = s * akp + c * akq; }
            for (long k = 0; k < n; k++) { const double apk = A(p, k), aqk = A(q, k); A(p, k) = c * apk - s * aqk; A(q, k) = s * apk + c * aqk; }
            for (long k = 0; k < n; k++) { const double vkp = V(k, p), vkq = V(k, q); V(k, p) = c * vkp - s * vkq; V(k, q) = s * vkp + c * vkq; }
        }
    }
    double m = 0; long idx = 0;
    for (long i = 0; i < n; i++) if (std::fabs(A(i, i)) > m) { m = std::fabs(A(i, i)); idx = i; }   // strict >, first wins
    Vec q(n);
    double nrm = 0; for (long i = 0; i < n; i++) { q[i] = V(i, idx); nrm += q[i] * q[i]; }
    nrm = std::sqrt(nrm);
    long big = 0; for (long i = 1; i < n; i++) if (std::fabs(q[i]) > std::fabs(q[big])) big = i;
    const double sgn = (q[big] < 0 ? -1.0 : 1.0) / nrm;
    for (auto& v : q) v *= sgn;
    return q;
}

// ---------------------------------------------------------------------------------------------
enum METHOD { KERNEL_TYPE1 = 0, KERNEL_TYPE2 = 1 };
enum VALIDATION_OUTPUT { RESS = 0, MSE = 1 };

// lib/PLS/include/PLS/pls.h:44-53 — Residual: M matrices (one per Y column), each n_obs x A
struct Residual { std::vector<Mat> E; std::string method; };

// lib/PLS/include/PLS/pls.h:184-266, lib/PLS/src/pls.cpp:340-510 — PLS::Model (real parts only; the
// reference stores complex<double> whose imaginary parts are zero for a real dominant eigenpair).
struct Model {
    Mat _X, _Y;
    size_t A;
    Mat P, W, R, Q, T;
    METHOD method;

    Model(const Mat& X, const Mat& Y, METHOD algorithm, size_t max_components)
        : _X(X), _Y(Y), A(max_components), method(algorithm) {                       // pls.cpp:340-353
        assert(max_components <= (size_t)_X.c);
        assert(_X.r != 0);
        assert(_X.r == _Y.r);
        P = Mat(_X.c, A); W = Mat(_X.c, A); R = Mat(_X.c, A); Q = Mat(_Y.c, A);
        plsr(_X, _Y, algorithm);
    }

    // pls.cpp:390-437 — Dayal & MacGregor (1997) modified kernel algorithms 1 and 2
    void plsr(const Mat& X, const Mat& Y, METHOD algorithm) {
        method = algorithm;
        const long N = X.r, K = X.c, M = Y.c;
        if (algorithm == KERNEL_TYPE1) T = Mat(N, A);
        Mat XY(K, M);                                                                 // :396  XY = X^T Y
        for (long m = 0; m < M; m++) for (long k = 0; k < K; k++) {
            const double* x = X.col(k); const double* y = Y.col(m);
            double s = 0; for (long i = 0; i < N; i++) s += x[i] * y[i];
            XY(k, m) = s;
        }
        Mat XX;
        if (algorithm == KERNEL_TYPE2) {                                              // :398  XX = X^T X
            XX = Mat(K, K);
            for (long a = 0; a < K; a++) for (long b = 0; b <= a; b++) {
                const double* xa = X.col(a); const double* xb = X.col(b);
                double s = 0; for (long i = 0; i < N; i++) s += xa[i] * xb[i];
                XX(a, b) = s; XX(b, a) = s;
            }
        }
        for (size_t i = 0; i < A; i++) {                                              // :400
            Vec w(K), p(K), q(M), r(K), t;
            double tt;
            if (M == 1) {                                                             // :403-404
                for (long k = 0; k < K; k++) w[k] = XY(k, 0);
            } else {                                                                  // :406-408
                Mat S(M, M);
                for (long a = 0; a < M; a++) for (long b = 0; b < M; b++) {
                    double s = 0; for (long k = 0; k < K; k++) s += XY(k, a) * XY(k, b);
                    S(a, b) = s;
                }
                q = dominant_eigenvector_sym(S);
                for (long k = 0; k < K; k++) { double s = 0; for (long m = 0; m < M; m++) s += XY(k, m) * q[m]; w[k] = s; }
            }
            { double ww = 0; for (long k = 0; k < K; k++) ww += w[k] * w[k]; ww = std::sqrt(ww); for (auto& v : w) v /= ww; }   // :411
            r = w;                                                                    // :412
            if (i != 0) for (size_t j = 0; j <= i - 1; j++) {                         // :414-416
                double pw = 0; for (long k = 0; k < K; k++) pw += P(k, j) * w[k];
                for (long k = 0; k < K; k++) r[k] -= pw * R(k, j);
            }
            if (algorithm == KERNEL_TYPE1) {                                          // :418-421
                t.assign(N, 0.0);
                for (long k = 0; k < K; k++) { const double* x = X.col(k); const double rk = r[k]; for (long n = 0; n < N; n++) t[n] += x[n] * rk; }
                tt = 0; for (long n = 0; n < N; n++) tt += t[n] * t[n];
                for (long k = 0; k < K; k++) { const double* x = X.col(k); double s = 0; for (long n = 0; n < N; n++) s += x[n] * t[n]; p[k] = s; }
            } else {                                                                  // :422-424
                Vec xr(K, 0.0);
                for (long b = 0; b < K; b++) { double s = 0; for (long a = 0; a < K; a++) s += r[a] * XX(a, b); xr[b] = s; }   // r^T XX
                tt = 0; for (long b = 0; b < K; b++) tt += xr[b] * r[b];
                p = xr;
            }
            for (auto& v : p) v /= tt;                                                // :427
            for (long m = 0; m < M; m++) { double s = 0; for (long k = 0; k < K; k++) s += r[k] * XY(k, m); q[m] = s / tt; }   // :428
            for (long m = 0; m < M; m++) for (long k = 0; k < K; k++) XY(k, m) -= (p[k] * q[m]) * tt;                          // :429
            for (long k = 0; k < K; k++) { W(k, i) = w[k]; P(k, i) = p[k]; R(k, i) = r[k]; }                                   // :430-433
            for (long m = 0; m < M; m++) Q(m, i) = q[m];
            if (algorithm == KERNEL_TYPE1) std::memcpy(T.col(i), t.data(), sizeof(double) * N);                               // :434
        }
    }

    // pls.cpp:439-442
    Mat scores(const Mat& X_new, size_t comp) const {
        assert(A >= comp);
        Mat out(X_new.r, comp);
        for (size_t a = 0; a < comp; a++) for (long k = 0; k < X_new.c; k++) {
            const double* x = X_new.col(k); double* o = out.col(a); const double rk = R(k, a);
            for (long n = 0; n < X_new.r; n++) o[n] += x[n] * rk;
        }
        return out;
    }
    // pls.cpp:444-447 — R[:, :comp] Q[:, :comp]^T  (K x M)
    Mat coefficients(size_t comp) const {
        assert(A >= comp);
        Mat B(R.r, Q.r);
        for (long m = 0; m < Q.r; m++) for (long k = 0; k < R.r; k++) {
            double s = 0; for (size_t a = 0; a < comp; a++) s += R(k, a) * Q(m, a);
            B(k, m) = s;
        }
        return B;
    }
    // pls.cpp:449-451
    Mat fitted_values(const Mat& X_new, size_t comp) const {
        const Mat B = coefficients(comp);
        Mat F(X_new.r, B.c);
        for (long m = 0; m < B.c; m++) for (long k = 0; k < X_new.c; k++) {
            const double* x = X_new.col(k); double* o = F.col(m); const double b = B(k, m);
            for (long n = 0; n < X_new.r; n++) o[n] += x[n] * b;
        }
        return F;
    }
    // pls.cpp:453-455
    Mat residuals(const Mat& X_new, const Mat& Y_new, size_t comp) const {
        Mat F = fitted_values(X_new, comp);
        for (size_t i = 0; i < F.d.size(); i++) F.d[i] = Y_new.d[i] - F.d[i];
        return F;
    }
    // pls.cpp:457-459
    Vec SSE(const Mat& X_new, const Mat& Y_new, size_t comp) const {
        const Mat E = residuals(X_new, Y_new, comp);
        Vec out(E.c, 0.0);
        for (long m = 0; m < E.c; m++) { const double* e = E.col(m); double s = 0; for (long n = 0; n < E.r; n++) s += e[n] * e[n]; out[m] = s; }
        return out;
    }
    // pls.cpp:461-467
    Vec explained_variance(const Mat& X_new, const Mat& Y_new, size_t comp) const {
        Vec sse = SSE(X_new, Y_new, comp), sst = SST(Y_new, colwise_mean(Y_new));
        for (size_t m = 0; m < sse.size(); m++) sse[m] = 1.0 - sse[m] / sst[m];
        return sse;
    }
    // pls.cpp:494-510
    Residual cv_NEW_DATA(const Mat& X_new, const Mat& Y_new) const {
        assert(X_new.c == _X.c && Y_new.c == _Y.c);
        Residual out; out.method = "NEW DATA";
        out.E.assign(Y_new.c, Mat(X_new.r, A));
        for (size_t nc = 1; nc <= A; nc++) {
            const Mat res = residuals(X_new, Y_new, nc);
            for (long y = 0; y < res.c; y++) std::memcpy(out.E[y].col(nc - 1), res.col(y), sizeof(double) * res.r);
        }
        return out;
    }
    // pls.cpp:469-491 — leave-one-out: the held-out slot walks 0..N-1 over a buffer that starts as rows 1..N-1
    Residual cv_LOO() const {
        const long N = _X.r;
        Mat Xv = _X.rows_range(1, N - 1), Yv = _Y.rows_range(1, N - 1);
        Residual out; out.method = "LOO";
        out.E.assign(_Y.c, Mat(N, A));
        Model plsm_v(Xv, Yv, method, (size_t)Xv.c);    // 2-arg public ctor at :477 -> max_components = X.cols()
        for (long row_out = 0; row_out < N; row_out++) {
            const Mat xr = _X.rows_range(row_out, 1), yr = _Y.rows_range(row_out, 1);
            for (size_t nc = 1; nc <= A; nc++) {
                const Mat res = plsm_v.residuals(xr, yr, nc);
                for (long k = 0; k < res.c; k++) out.E[k](row_out, nc - 1) = res(0, k);
            }
            if (row_out < Xv.r) {
                for (long k = 0; k < Xv.c; k++) Xv(row_out, k) = _X(row_out, k);
                for (long k = 0; k < Yv.c; k++) Yv(row_out, k) = _Y(row_out, k);
                plsm_v.plsr(Xv, Yv, method);
            }
        }
        return out;
    }
    // pls.cpp:512-549 — leave-some-out. `shuffles` holds, per trial, the index vector `full` as rand_nchoosek (:218-227) leaves it
    // after its std::shuffle (libstdc++-specific, so it stays with the caller): full[0, train) trains, full[train, N) is predicted.
    Residual cv_LSO(const uint64_t* shuffles, long test_size, long num_trials) const {
        const long N = _X.r, train_size = N - test_size;
        assert(test_size != 0 && train_size != 0);
        Residual out; out.method = "LSO";
        out.E.assign(_Y.c, Mat(num_trials * test_size, A));
        for (long rep = 0; rep < num_trials; rep++) {
            const uint64_t* full = shuffles + (size_t)rep * N;
            Mat Xv(train_size, _X.c), Yv(train_size, _Y.c), Xp(test_size, _X.c), Yp(test_size, _Y.c);
            for (long i = 0; i < N; i++) {
                const long src = (long)full[i];
                if (i < train_size) { for (long k = 0; k < _X.c; k++) Xv(i, k) = _X(src, k); for (long k = 0; k < _Y.c; k++) Yv(i, k) = _Y(src, k); }
                else { for (long k = 0; k < _X.c; k++) Xp(i - train_size, k) = _X(src, k); for (long k = 0; k < _Y.c; k++) Yp(i - train_size, k) = _Y(src, k); }
            }
            Model plsm_v(Xv, Yv, method, (size_t)Xv.c);                                // :539 (A' = num_predictors, :529)
            for (size_t nc = 1; nc <= A; nc++) {
                const Mat res = plsm_v.residuals(Xp, Yp, nc);
                for (long y = 0; y < res.c; y++) for (long i = 0; i < test_size; i++) out.E[y](rep * test_size + i, nc - 1) += res(i, y);   // :543
            }
        }
        return out;
    }
};

// lib/PLS/src/pls.cpp:235-261 — validation(): rows = Y component, cols = #components
static Mat validation(const Residual& res, VALIDATION_OUTPUT out_type) {
    if (res.E.empty()) return Mat(0, 0);
    Mat SSEv((long)res.E.size(), res.E[0].c);
    for (size_t y = 0; y < res.E.size(); y++) for (long c = 0; c < res.E[y].c; c++) {
        const double* e = res.E[y].col(c); double s = 0;
        for (long n = 0; n < res.E[y].r; n++) s += e[n] * e[n];
        SSEv((long)y, c) += s;
    }
    if (out_type == MSE) { const double n = (double)res.E[0].r; for (auto& v : SSEv.d) v /= n; }
    return SSEv;
}

// lib/PLS/src/pls.cpp:265-289 — optimal_num_components(): first argmin PRESS, then the smallest
// alt < ref whose Wilcoxon p-value against ref exceeds ALPHA; returned as component COUNTS (index+1)
static std::vector<size_t> optimal_num_components(const Residual& res, double ALPHA, Mat* press_out = nullptr) {
    const Mat press = validation(res, RESS);
    if (press_out) *press_out = press;
    std::vector<size_t> min_press_idx(press.r);
    for (size_t y = 0; y < res.E.size(); y++) {
        size_t best = 0;                                         // Eigen minCoeff(&idx): first minimum
        for (long c = 1; c < press.c; c++) if (press((long)y, c) < press((long)y, (long)best)) best = (size_t)c;
        min_press_idx[y] = best;
        const size_t ref_min = best;
        const double* err1 = res.E[y].col((long)ref_min);
        for (size_t alt = 0; alt < ref_min; alt++) {
            const double* err2 = res.E[y].col((long)alt);
            if (wilcoxon(err1, err2, (size_t)res.E[y].r) > ALPHA) { min_press_idx[y] = alt; break; }
        }
    }
    for (auto& v : min_press_idx) v += 1;
    return min_press_idx;
}

// src/AbcUtil.cpp:320-324 — euclidean(): rowwise norm of (sims - ref)
static Vec euclidean(const Mat& sims, const Vec& ref) {
    Vec acc(sims.r, 0.0);
    for (long k = 0; k < sims.c; k++) { const double* x = sims.col(k); for (long n = 0; n < sims.r; n++) { const double dlt = x[n] - ref[k]; acc[n] += dlt * dlt; } }
    for (auto& v : acc) v = std::sqrt(v);
    return acc;
}

// include/AbcSmc/RunningStat.h:16-46 — Welford; src/AbcUtil.cpp:528-537 — 2 x sample variance per column
static Vec calculate_doubled_variance(const Mat& params) {
    Vec v2(params.c, 0.0);
    for (long p = 0; p < params.c; p++) {
        int m_n = 0; double oldM = 0, newM = 0, oldS = 0, newS = 0;
        const double* xs = params.col(p);
        for (long i = 0; i < params.r; i++) {
            const double x = xs[i];
            m_n++;
            if (m_n == 1) { oldM = newM = x; oldS = 0.0; }
            else { newM = oldM + (x - oldM) / m_n; newS = oldS + (x - oldM) * (x - newM); oldM = newM; oldS = newS; }
        }
        v2[p] = 2 * ((m_n > 1) ? newS / (m_n - 1) : 0.0);
    }
    return v2;
}

// GSL randist/gauss.c gsl_ran_gaussian_pdf (see header)
static inline double gsl_ran_gaussian_pdf(const double x, const double sigma) {
    const double u = x / std::fabs(sigma);
    return (1 / (std::sqrt(2 * M_PI) * std::fabs(sigma))) * std::exp(-u * u / 2);
}

// include/AbcSmc/Priors.h:53-55 (Gaussian), :75-77 (discrete uniform), :101-103 (continuous uniform)
enum PRIOR { PRIOR_UNIFORM = 0, PRIOR_DISCRETE_UNIFORM = 1, PRIOR_GAUSSIAN = 2 };
static double prior_likelihood(int type, double a, double b, double pval) {
    switch (type) {
        case PRIOR_UNIFORM: return ((a <= pval) && (pval <= b)) ? 1.0 / (b - a) : 0.0;
        case PRIOR_DISCRETE_UNIFORM: { const long mn = (long)a, mx = (long)b;
            return ((pval == std::round(pval)) && (mn <= pval) && (pval <= mx)) ? 1.0 / (mx - mn + 1) : 0.0; }
        default: return gsl_ran_gaussian_pdf(pval - a, b);   // a = mean, b = sd
    }
}

// ---- next-set proposal sampling -------------------------------------------------------------------
// src/AbcUtil.cpp:111-121 (gsl_rng_nonuniform_int), :146-158 (gsl_ran_trunc_normal), :366-390 (sample_posterior,
// sample_predictive_priors); include/AbcSmc/Priors.h:18-41 (Prior::noise / trynoise), :80 (recast of the discrete prior).
// The reference draws from a gsl_rng (absent here). This restatement keeps the reference's ORDER of draws — all parent rows
// first (sample_posterior is evaluated before the noise loop, :384), then row by row, parameter by parameter, attempt by
// attempt — on a splitmix64 stream: gsl_ran_discrete (Walker alias) is restated by inverting the cumulative weights (the same
// distribution P(j) = w_j / sum w), gsl_ran_gaussian by the Marsaglia polar method GSL itself uses (randist/gauss.c).
// It is the DISTRIBUTIONAL checker for abcb200_sample_predictive_priors, not a bit-level one.
struct SplitMix64 {
    uint64_t s;
    uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
    double uniform() { return ((double)(next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }            // (0, 1)
    double gaussian(double sigma) {                                                                       // polar (Marsaglia)
        double x, y, r2;
        do { x = -1.0 + 2.0 * uniform(); y = -1.0 + 2.0 * uniform(); r2 = x * x + y * y; } while (r2 > 1.0 || r2 == 0.0);
        return sigma * y * std::sqrt(-2.0 * std::log(r2) / r2);
    }
};
static double prior_recast(int type, double v) { return type == PRIOR_DISCRETE_UNIFORM ? std::round(v) : v; }   // Priors.h:57, 80, 105
static double prior_mean(int type, double a, double b) { return type == PRIOR_GAUSSIAN ? a : (a + b) / 2.0; }      // Priors.h:66, 92; Gaussian: meanval
// Prior::noise (Priors.h:18-33)
static double prior_noise(SplitMix64& rng, int type, double a, double b, double mu, double sigma, long max_attempts, long* fallbacks) {
    long attempts = 1;
    double dev = prior_recast(type, rng.gaussian(sigma) + mu);
    while (!(prior_likelihood(type, a, b, dev) != 0.0) && (attempts++ < max_attempts)) dev = prior_recast(type, rng.gaussian(sigma) + mu);
    if (!(prior_likelihood(type, a, b, dev) != 0.0)) { if (fallbacks) (*fallbacks)++; return prior_mean(type, a, b); }
    return dev;
}
static void sample_predictive_priors(uint64_t seed, long num_samples, const Vec& weights, const Mat& parameter_prior, const int* ptype,
                                     const double* pa, const double* pb, const Vec& doubled_variance, long max_attempts, Mat& noised,
                                     std::vector<uint64_t>& parent, long* fallbacks) {
    SplitMix64 rng{seed};
    const long n = parameter_prior.r, P = parameter_prior.c;
    std::vector<double> cdf(n);
    double acc = 0.0;
    for (long j = 0; j < n; j++) { acc += weights[j]; cdf[j] = acc; }
    parent.resize(num_samples);
    for (long i = 0; i < num_samples; i++) {                       // AbcUtil.cpp:117
        const double x = rng.uniform() * acc;
        parent[i] = (uint64_t)(std::upper_bound(cdf.begin(), cdf.end(), x) - cdf.begin());
        if (parent[i] >= (uint64_t)n) parent[i] = n - 1;
    }
    for (long i = 0; i < num_samples; i++)                         // AbcUtil.cpp:386-388
        for (long p = 0; p < P; p++)                               // AbcUtil.cpp:152-156
            noised(i, p) = prior_noise(rng, ptype[p], pa[p], pb[p], parameter_prior((long)parent[i], p), std::sqrt(doubled_variance[p]), max_attempts, fallbacks);
}

// src/AbcUtil.cpp:462-488 setup_mvn_sampler: gsl_ran_multivariate_gaussian_vcov (GSL manual: sum (x_i - mu)(x_i - mu)^T / (n - 1)),
// diagonal doubled (:477-480), gsl_linalg_cholesky_decomp1 (lower factor L; GSL_EDOM when not positive definite -> returns false).
static bool setup_mvn_sampler(const Mat& params, Mat& L) {
    const long n = params.r, P = params.c;
    Vec mu(P, 0.0);
    for (long p = 0; p < P; p++) { double s = 0.0; for (long i = 0; i < n; i++) s += params(i, p); mu[p] = s / (double)n; }
    L = Mat(P, P);
    for (long j = 0; j < P; j++)
        for (long k = 0; k <= j; k++) {
            double s = 0.0;
            for (long i = 0; i < n; i++) s += (params(i, j) - mu[j]) * (params(i, k) - mu[k]);
            s /= (double)(n - 1);
            if (j == k) s *= 2.0;
            L(j, k) = s; L(k, j) = s;
        }
    for (long c = 0; c < P; c++) {                    // Cholesky, lower, column by column
        double d = L(c, c);
        for (long q = 0; q < c; q++) d -= L(c, q) * L(c, q);
        if (!(d > 0.0)) return false;
        d = std::sqrt(d);
        L(c, c) = d;
        for (long i = c + 1; i < P; i++) { double v = L(i, c); for (long q = 0; q < c; q++) v -= L(i, q) * L(c, q); L(i, c) = v / d; }
    }
    for (long j = 0; j < P; j++) for (long i = 0; i < j; i++) L(i, j) = 0.0;
    return true;
}
// src/AbcUtil.cpp:392-404 sample_mvn_predictive_priors + :123-144 gsl_ran_trunc_mv_normal (same stream conventions as above;
// the reference has no attempt limit — max_attempts only protects the test harness)
static void sample_mvn_predictive_priors(uint64_t seed, long num_samples, const Vec& weights, const Mat& parameter_prior, const int* ptype,
                                         const double* pa, const double* pb, const Mat& L, long max_attempts, Mat& noised,
                                         std::vector<uint64_t>& parent, long* failures) {
    SplitMix64 rng{seed};
    const long n = parameter_prior.r, P = parameter_prior.c;
    std::vector<double> cdf(n);
    double acc = 0.0;
    for (long j = 0; j < n; j++) { acc += weights[j]; cdf[j] = acc; }
    parent.resize(num_samples);
    for (long i = 0; i < num_samples; i++) {
        const double x = rng.uniform() * acc;
        parent[i] = (uint64_t)(std::upper_bound(cdf.begin(), cdf.end(), x) - cdf.begin());
        if (parent[i] >= (uint64_t)n) parent[i] = n - 1;
    }
    Vec z(P), res(P);
    for (long i = 0; i < num_samples; i++) {
        bool success = false;
        long attempts = 0;
        while (!success && attempts++ < max_attempts) {
            success = true;
            for (long p = 0; p < P; p++) z[p] = rng.gaussian(1.0);                       // gsl_ran_multivariate_gaussian: mu + L z
            for (long p = 0; success && p < P; p++) {
                double v = parameter_prior((long)parent[i], p);
                for (long q = 0; q <= p; q++) v += L(p, q) * z[q];
                res[p] = prior_recast(ptype[p], v);
                success = prior_likelihood(ptype[p], pa[p], pb[p], res[p]) != 0.0;
            }
        }
        if (!success) { if (failures) (*failures)++; for (long p = 0; p < P; p++) res[p] = prior_recast(ptype[p], parameter_prior((long)parent[i], p)); }
        for (long p = 0; p < P; p++) noised(i, p) = res[p];
    }
}

// src/AbcUtil.cpp:547-586 — SMC importance weights for set t > 0, L2-normalised (Eigen normalize())
static Vec weight_predictive_prior(const Vec& numer, const Mat& params, const Mat& prev_params,
                                   const Vec& prev_weights, const Vec& prev_dv) {
    Vec weight(params.r, 0.0);
    for (long i = 0; i < params.r; i++) {
        const double numerator = numer[i];
        double denominator = 0.0;
        for (long j = 0; j < prev_params.r; j++) {
            double running_product = prev_weights[j];
            for (long p = 0; p < prev_params.c; p++) {
                const double par_value = params(i, p), old_par_value = prev_params(j, p), old_dv = prev_dv[p];
                if (old_dv != 0 || par_value != old_par_value)
                    running_product *= gsl_ran_gaussian_pdf(par_value - old_par_value, std::sqrt(old_dv));
            }
            denominator += running_product;
        }
        weight[i] = numerator / denominator;
    }
    double n2 = 0; for (double w : weight) n2 += w * w;    // Eigen normalize(): divide by the 2-norm
    if (n2 > 0) { n2 = std::sqrt(n2); for (auto& w : weight) w /= n2; }   // Eigen: if (squaredNorm > 0) v /= sqrt(.)
    return weight;
}

struct RankResult { std::vector<size_t> order; Vec dist; std::vector<size_t> ncomp; size_t ncomp_used; Mat press; };

// src/AbcUtil.cpp:423-458 — particle_ranking_PLS
static RankResult particle_ranking_PLS(const Mat& metric_vals, const Mat& param_vals, const Vec& target, double training_fraction) {
    assert((0 < training_fraction) && (training_fraction <= 1));
    const Vec met_means = colwise_mean(metric_vals);
    const Vec met_stdev = colwise_stdev(metric_vals, met_means);
    const Mat z_met = colwise_z_scores(metric_vals, met_means, met_stdev);
    const Vec par_means = colwise_mean(param_vals);
    const Mat z_par = colwise_z_scores(param_vals, par_means, colwise_stdev(param_vals, par_means));
    const Vec obs_met = z_scores(target, met_means, met_stdev);
    const size_t n_tr = (size_t)std::round(z_met.r * training_fraction);
    Model plsm(z_met.rows_range(0, (long)n_tr), z_par.rows_range(0, (long)n_tr), KERNEL_TYPE1, (size_t)z_met.c);
    const size_t n_te = z_met.r - n_tr;
    RankResult out;
    {
        const Residual em = plsm.cv_NEW_DATA(z_met.rows_range((long)n_tr, (long)n_te), z_par.rows_range((long)n_tr, (long)n_te));
        out.ncomp = optimal_num_components(em, 0.1, &out.press);
    }
    out.ncomp_used = *std::max_element(out.ncomp.begin(), out.ncomp.end());
    Mat obs_m(1, z_met.c); for (long k = 0; k < z_met.c; k++) obs_m(0, k) = obs_met[k];
    const Mat obs_s = plsm.scores(obs_m, out.ncomp_used);
    Vec obs_scores(out.ncomp_used); for (size_t a = 0; a < out.ncomp_used; a++) obs_scores[a] = obs_s(0, (long)a);
    const Mat sim_scores = plsm.scores(z_met, out.ncomp_used);
    out.dist = euclidean(sim_scores, obs_scores);
    out.order = ordered(out.dist.data(), out.dist.size());
    return out;
}

// src/AbcUtil.cpp:408-421 — particle_ranking_simple
static RankResult particle_ranking_simple(const Mat& X_orig, const Vec& target) {
    const Vec mu = colwise_mean(X_orig);
    const Vec sd = colwise_stdev(X_orig, mu);
    const Vec obs = z_scores(target, mu, sd);
    const Mat X = colwise_z_scores(X_orig, mu, sd);
    RankResult out;
    out.dist = euclidean(X, obs);
    out.order = ordered(out.dist.data(), out.dist.size());
    out.ncomp_used = 0;
    return out;
}

}  // namespace orc

// =============================================================================================
// C interface for ctypes (tests/, bench.py cpu_baseline). Column-major, ld == rows.
// =============================================================================================
using namespace orc;
static Vec to_vec(const double* p, long n) { return Vec(p, p + n); }

extern "C" {

void orc_colwise_mean(const double* X, long n, long k, double* out) { Vec v = colwise_mean(Mat(n, k, X, n)); std::copy(v.begin(), v.end(), out); }
void orc_colwise_stdev(const double* X, long n, long k, const double* mean, double* out) { Vec v = colwise_stdev(Mat(n, k, X, n), to_vec(mean, k)); std::copy(v.begin(), v.end(), out); }
void orc_colwise_z_scores(const double* X, long n, long k, const double* mean, const double* sd, double* Z) {
    Mat z = colwise_z_scores(Mat(n, k, X, n), to_vec(mean, k), to_vec(sd, k)); std::copy(z.d.begin(), z.d.end(), Z);
}
void orc_colwise_z_scores_auto(const double* X, long n, long k, double* Z) {     // pls.cpp:107-111
    Mat m(n, k, X, n); Vec mu = colwise_mean(m); Vec sd = colwise_stdev(m, mu);
    Mat z = colwise_z_scores(m, mu, sd); std::copy(z.d.begin(), z.d.end(), Z);
}
void orc_z_scores(const double* obs, const double* mean, const double* sd, long k, double* out) { Vec v = z_scores(to_vec(obs, k), to_vec(mean, k), to_vec(sd, k)); std::copy(v.begin(), v.end(), out); }
double orc_normalcdf(double z) { return normalcdf(z); }
double orc_wilcoxon(const double* e1, const double* e2, long n) { return wilcoxon(e1, e2, (size_t)n); }
void orc_ordered(const double* v, long n, uint64_t* out) { auto o = ordered(v, (size_t)n); for (long i = 0; i < n; i++) out[i] = o[i]; }
void orc_euclidean(const double* S, long n, long k, const double* ref, double* out) { Vec v = euclidean(Mat(n, k, S, n), to_vec(ref, k)); std::copy(v.begin(), v.end(), out); }
void orc_dominant_eigenvector_sym(const double* S, long n, double* out) { Vec v = dominant_eigenvector_sym(Mat(n, n, S, n)); std::copy(v.begin(), v.end(), out); }

void* orc_pls_fit(const double* X, const double* Y, long n, long K, long M, int method, long max_components) {
    return new Model(Mat(n, K, X, n), Mat(n, M, Y, n), (METHOD)method, (size_t)max_components);
}
void orc_pls_free(void* m) { delete (Model*)m; }
// which: 'P','W','R' (K x A), 'Q' (M x A), 'T' (N x A; type 1 only)
void orc_pls_get(void* mp, char which, double* out) {
    Model* m = (Model*)mp; const Mat* s = nullptr;
    switch (which) { case 'P': s = &m->P; break; case 'W': s = &m->W; break; case 'R': s = &m->R; break; case 'Q': s = &m->Q; break; default: s = &m->T; }
    std::copy(s->d.begin(), s->d.end(), out);
}
void orc_pls_scores(void* mp, const double* Xn, long n, long comp, double* out) { Model* m = (Model*)mp; Mat s = m->scores(Mat(n, m->_X.c, Xn, n), (size_t)comp); std::copy(s.d.begin(), s.d.end(), out); }
void orc_pls_coefficients(void* mp, long comp, double* out) { Mat b = ((Model*)mp)->coefficients((size_t)comp); std::copy(b.d.begin(), b.d.end(), out); }
void orc_pls_fitted_values(void* mp, const double* Xn, long n, long comp, double* out) { Model* m = (Model*)mp; Mat f = m->fitted_values(Mat(n, m->_X.c, Xn, n), (size_t)comp); std::copy(f.d.begin(), f.d.end(), out); }
void orc_pls_residuals(void* mp, const double* Xn, const double* Yn, long n, long comp, double* out) {
    Model* m = (Model*)mp; Mat e = m->residuals(Mat(n, m->_X.c, Xn, n), Mat(n, m->_Y.c, Yn, n), (size_t)comp); std::copy(e.d.begin(), e.d.end(), out);
}
void orc_pls_SSE(void* mp, const double* Xn, const double* Yn, long n, long comp, double* out) {
    Model* m = (Model*)mp; Vec v = m->SSE(Mat(n, m->_X.c, Xn, n), Mat(n, m->_Y.c, Yn, n), (size_t)comp); std::copy(v.begin(), v.end(), out);
}
void orc_pls_explained_variance(void* mp, const double* Xn, const double* Yn, long n, long comp, double* out) {
    Model* m = (Model*)mp; Vec v = m->explained_variance(Mat(n, m->_X.c, Xn, n), Mat(n, m->_Y.c, Yn, n), (size_t)comp); std::copy(v.begin(), v.end(), out);
}
void* orc_pls_cv_new_data(void* mp, const double* Xn, const double* Yn, long n) {
    Model* m = (Model*)mp; return new Residual(m->cv_NEW_DATA(Mat(n, m->_X.c, Xn, n), Mat(n, m->_Y.c, Yn, n)));
}
void* orc_pls_cv_loo(void* mp) { return new Residual(((Model*)mp)->cv_LOO()); }
void* orc_pls_cv_lso(void* mp, const uint64_t* shuffles, long test_size, long num_trials) { return new Residual(((Model*)mp)->cv_LSO(shuffles, test_size, num_trials)); }
void orc_residual_free(void* r) { delete (Residual*)r; }
long orc_residual_rows(void* r) { return ((Residual*)r)->E.empty() ? 0 : ((Residual*)r)->E[0].r; }
// errors cube out: [y][c][n] contiguous (M matrices, each column-major n x A)
void orc_residual_errors(void* rp, double* out) { Residual* r = (Residual*)rp; size_t off = 0; for (auto& e : r->E) { std::copy(e.d.begin(), e.d.end(), out + off); off += e.d.size(); } }
void orc_validation(void* rp, int out_type, double* out /* M x A col-major */) { Mat v = validation(*(Residual*)rp, (VALIDATION_OUTPUT)out_type); std::copy(v.d.begin(), v.d.end(), out); }
void orc_optimal_num_components(void* rp, double alpha, uint64_t* out) { auto v = optimal_num_components(*(Residual*)rp, alpha); for (size_t i = 0; i < v.size(); i++) out[i] = v[i]; }

void orc_particle_ranking_PLS(const double* met, const double* par, long N, long K, long P, const double* target,
                              double training_fraction, uint64_t* order_out, double* dist_out, uint64_t* ncomp_out,
                              uint64_t* ncomp_used_out, double* press_out) {
    RankResult r = particle_ranking_PLS(Mat(N, K, met, N), Mat(N, P, par, N), to_vec(target, K), training_fraction);
    for (long i = 0; i < N; i++) order_out[i] = r.order[i];
    if (dist_out) std::copy(r.dist.begin(), r.dist.end(), dist_out);
    if (ncomp_out) for (long p = 0; p < P; p++) ncomp_out[p] = r.ncomp[p];
    if (ncomp_used_out) *ncomp_used_out = r.ncomp_used;
    if (press_out) std::copy(r.press.d.begin(), r.press.d.end(), press_out);
}
void orc_particle_ranking_simple(const double* met, long N, long K, const double* target, uint64_t* order_out, double* dist_out) {
    RankResult r = particle_ranking_simple(Mat(N, K, met, N), to_vec(target, K));
    for (long i = 0; i < N; i++) order_out[i] = r.order[i];
    if (dist_out) std::copy(r.dist.begin(), r.dist.end(), dist_out);
}
void orc_calculate_doubled_variance(const double* params, long n, long P, double* out) { Vec v = calculate_doubled_variance(Mat(n, P, params, n)); std::copy(v.begin(), v.end(), out); }
double orc_gsl_ran_gaussian_pdf(double x, double sigma) { return gsl_ran_gaussian_pdf(x, sigma); }
double orc_prior_likelihood(int type, double a, double b, double v) { return prior_likelihood(type, a, b, v); }
// src/AbcUtil.cpp:539-545 — set 0: uniform 1/N, not normalised
void orc_weight_predictive_prior0(long n, double* out) { const double u = 1.0 / (double)n; for (long i = 0; i < n; i++) out[i] = u; }
// numer[i] = prod_p prior_p.likelihood(theta_new[i,p]) (src/AbcUtil.cpp:559-561), computed by the caller
void orc_weight_predictive_prior(const double* numer, const double* params, long n_new, const double* prev_params, long n_old,
                                 const double* prev_w, const double* prev_dv, long P, double* out) {
    Vec w = weight_predictive_prior(to_vec(numer, n_new), Mat(n_new, P, params, n_new), Mat(n_old, P, prev_params, n_old), to_vec(prev_w, n_old), to_vec(prev_dv, P));
    std::copy(w.begin(), w.end(), out);
}

// next-set proposals: ptype/pa/pb as in orc_prior_likelihood (uniform: [a, b]; discrete uniform: [a, b]; Gaussian: mean a, sd b)
void orc_sample_predictive_priors(uint64_t seed, long num_samples, const double* weights, const double* theta, long n_pp, long P, const int* ptype,
                                  const double* pa, const double* pb, const double* dv, long max_attempts, double* out, uint64_t* parent_out, long* fallbacks_out) {
    Mat noised(num_samples, P);
    std::vector<uint64_t> parent;
    long fb = 0;
    sample_predictive_priors(seed, num_samples, to_vec(weights, n_pp), Mat(n_pp, P, theta, n_pp), ptype, pa, pb, to_vec(dv, P), max_attempts, noised, parent, &fb);
    std::copy(noised.d.begin(), noised.d.end(), out);
    if (parent_out) std::copy(parent.begin(), parent.end(), parent_out);
    if (fallbacks_out) *fallbacks_out = fb;
}

int orc_setup_mvn_sampler(const double* theta, long n_pp, long P, double* L_out) {
    Mat L;
    if (!setup_mvn_sampler(Mat(n_pp, P, theta, n_pp), L)) return 1;
    std::copy(L.d.begin(), L.d.end(), L_out);
    return 0;
}
void orc_sample_mvn_predictive_priors(uint64_t seed, long num_samples, const double* weights, const double* theta, long n_pp, long P, const int* ptype,
                                      const double* pa, const double* pb, const double* L, long max_attempts, double* out, uint64_t* parent_out,
                                      long* failures_out) {
    Mat noised(num_samples, P);
    std::vector<uint64_t> parent;
    long fl = 0;
    sample_mvn_predictive_priors(seed, num_samples, to_vec(weights, n_pp), Mat(n_pp, P, theta, n_pp), ptype, pa, pb, Mat(P, P, L, P), max_attempts, noised, parent, &fl);
    std::copy(noised.d.begin(), noised.d.end(), out);
    if (parent_out) std::copy(parent.begin(), parent.end(), parent_out);
    if (failures_out) *failures_out = fl;
}

}  // extern "C"
