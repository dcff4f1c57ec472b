// flatten_test.cpp — host logic of the adapter, no GPU: ABC_B200::flatten_prior recovers what Prior::noise needs (validity interval,
// rounding recast, prior mean) from the reference's Parameter interface (include/AbcSmc/Parameter.h:51-77, Priors.h), which has no
// accessor for the bounds. The three prior kinds of the reference are restated minimally here (Priors.h:44-110).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>

#include "../../abcsmc_b200/host/abc_b200.hpp"

struct Prior {
    virtual ~Prior() {}
    virtual double likelihood(double v) const = 0;
    virtual double recast(double v) const { return v; }
    bool valid(double v) const { return likelihood(v) != 0.0; }                 // Parameter.h:77
    virtual double get_mean() const = 0;
    virtual double get_sd() const = 0;
};
struct ContinuousUniform : Prior {                                              // Priors.h:86-110
    ContinuousUniform(double a, double b) : a_(a), b_(b) {}
    double likelihood(double v) const override { return (a_ <= v && v <= b_) ? 1.0 / (b_ - a_) : 0.0; }
    double get_mean() const override { return (b_ + a_) / 2.0; }
    double get_sd() const override { return (b_ - a_) / std::sqrt(12.0); }
    double a_, b_;
};
struct DiscreteUniform : Prior {                                                // Priors.h:60-84
    DiscreteUniform(long a, long b) : a_(a), b_(b) {}
    double recast(double v) const override { return std::round(v); }
    double likelihood(double v) const override { return (v == recast(v) && a_ <= v && v <= b_) ? 1.0 / (double)(b_ - a_ + 1) : 0.0; }
    double get_mean() const override { return (double)(b_ + a_) / 2.0; }
    double get_sd() const override { return (double)(b_ - a_) / std::sqrt(12.0); }
    long a_, b_;
};
struct Gaussian : Prior {                                                       // Priors.h:44-58
    Gaussian(double m, double s) : m_(m), s_(s) {}
    double likelihood(double v) const override { const double u = (v - m_) / std::fabs(s_); return std::exp(-u * u / 2.0) / (std::sqrt(2.0 * M_PI) * std::fabs(s_)); }
    double get_mean() const override { return m_; }
    double get_sd() const override { return s_; }
    double m_, s_;
};

static int fails = 0;
#define CHECK(cond) do { if (!(cond)) { std::fprintf(stderr, "FAIL line %d: %s\n", __LINE__, #cond); fails++; } } while (0)

int main() {
    const double inf = std::numeric_limits<double>::infinity();
    const double bounds[][2] = {{0.0, 1.0}, {0.0, 2.0}, {-3.5, 7.25}, {1e-3, 1e-2}, {0.1, 0.3}, {-1e6, 2e6}, {0.3333333333333333, 0.7}};
    for (auto& b : bounds) {
        const ContinuousUniform p(b[0], b[1]);
        const ABC_B200::FlatPrior f = ABC_B200::flatten_prior(p);
        CHECK(f.lo == b[0]); CHECK(f.hi == b[1]); CHECK(f.integral == 0); CHECK(f.mean == (b[0] + b[1]) / 2.0);
        CHECK(p.valid(f.lo) && !p.valid(std::nextafter(f.lo, -inf)));
        CHECK(p.valid(f.hi) && !p.valid(std::nextafter(f.hi, inf)));
    }
    const long ib[][2] = {{0, 10}, {1, 6}, {-4, 3}, {0, 1}, {100, 100000}};
    for (auto& b : ib) {
        const DiscreteUniform p(b[0], b[1]);
        const ABC_B200::FlatPrior f = ABC_B200::flatten_prior(p);
        CHECK(f.lo == (double)b[0]); CHECK(f.hi == (double)b[1]); CHECK(f.integral == 1); CHECK(f.mean == (double)(b[0] + b[1]) / 2.0);
    }
    {
        const Gaussian p(1.0, 2.0);
        const ABC_B200::FlatPrior f = ABC_B200::flatten_prior(p);
        CHECK(f.lo == -inf); CHECK(f.hi == inf); CHECK(f.integral == 0); CHECK(f.mean == 1.0);
    }
    if (fails) return 1;
    std::printf("flatten ok\n");
    return 0;
}
