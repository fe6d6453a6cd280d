// CPU-only: accuracy of include/fastad_b200/fastmath.cuh (host build of the same header that the
// Black-Scholes kernel compiles for the device) against glibc.
#include <random>

#include <fastad_b200/fastmath.cuh>
#include "check.hpp"

int main()
{
    std::mt19937_64 gen(12345);
    std::uniform_real_distribution<double> U(0., 1.);
    double e_logt = 0, e_logt1 = 0, e_expt = 0, e_exp = 0, e_log = 0, e_div = 0, e_sqrt = 0, e_phi_abs = 0, e_phi_tail = 0, e_E = 0, e_pw_abs = 0, e_pw_tail = 0, e_pw_vs = 0;
    for (int i = 0; i < 2000000; ++i) {
        const double a = U(gen), b = U(gen);
        const double x = -60.0 * a * a + 2.0 * (b - 0.5);                 // exp arguments, mostly negative
        e_exp = std::max(e_exp, ulp_diff(fm::exp_(x), std::exp(x)));
        e_expt = std::max(e_expt, ulp_diff(fm::exp_t(x, fm::bs_table_host() + 256), std::exp(x)));
        const double y = std::exp(8.0 * (a - 0.5));                       // log arguments 0.018 .. 55
        e_log = std::max(e_log, ulp_diff(fm::log_(y), std::log(y)));
        e_logt = std::max(e_logt, ulp_diff(fm::log_t(y, fm::bs_table_host() + 288), std::log(y)));
        const double y1 = 1.0 + (a - 0.5) * std::ldexp(1.0, -(int)(40 * b));   // arguments approaching 1 from both sides
        e_logt1 = std::max(e_logt1, ulp_diff(fm::log_t(y1, fm::bs_table_host() + 288), std::log(y1)));
        const double p = 0.01 + 300 * a, q = 0.01 + 300 * b;
        e_div = std::max(e_div, ulp_diff(fm::div_(p, q), p / q));
        e_sqrt = std::max(e_sqrt, ulp_diff(fm::sqrt_(p), std::sqrt(p)));
        const double d = 16.0 * (a - 0.5);                                // d1/d2 range
        double Phi, PhiN, E;
        fm::phi_pair(d, Phi, PhiN, E);
        const double u = d * 0.70710678118654752440;
        const double ref = 0.5 * std::erfc(-u), refn = 0.5 * std::erfc(u);
        e_phi_abs = std::max(e_phi_abs, std::max(std::fabs(Phi - ref), std::fabs(PhiN - refn)));
        const double tail = d < 0 ? ulp_diff(Phi, ref) : ulp_diff(PhiN, refn);
        e_phi_tail = std::max(e_phi_tail, tail);
        e_E = std::max(e_E, ulp_diff(E, std::exp(-(u * u))));
        // piecewise-table variant (what the Black-Scholes kernel uses)
        double Phi2, PhiN2, E2;
        fm::phi_pair_pw(d, fm::bs_table_host(), Phi2, PhiN2, E2);
        e_pw_abs = std::max(e_pw_abs, std::max(std::fabs(Phi2 - ref), std::fabs(PhiN2 - refn)));
        e_pw_tail = std::max(e_pw_tail, d < 0 ? ulp_diff(Phi2, ref) : ulp_diff(PhiN2, refn));
        e_pw_vs = std::max(e_pw_vs, std::max(std::fabs(Phi2 - Phi), std::fabs(PhiN2 - PhiN)));
        e_E = std::max(e_E, ulp_diff(E2, std::exp(-(u * u))));   // the table exp: same 1-ulp bound
    }
    std::printf("piecewise Phi: abs %.3g, tail ulp %.2f, vs single polynomial abs %.3g\n", e_pw_abs, e_pw_tail, e_pw_vs);
    CHECK(e_pw_abs <= 4e-16);
    CHECK(e_pw_tail <= 8.0);
    std::printf("max ulp: exp %.2f log %.2f div %.2f sqrt %.2f gauss %.2f | Phi abs %.3g, tail ulp %.2f\n",
                e_exp, e_log, e_div, e_sqrt, e_E, e_phi_abs, e_phi_tail);
    std::printf("table exp: max ulp %.2f\n", e_expt);
    CHECK(e_exp <= 1.0);
    std::printf("table log: max ulp %.2f, near 1: %.2f\n", e_logt, e_logt1);
    CHECK(e_logt <= 1.0);
    CHECK(e_logt1 <= 1.0);
    CHECK(fm::log_t(1.0, fm::bs_table_host() + 288) == 0.0);
    for (double xx : {0.75, 1.5, 0.7499999999999999, 1.4999999999999998, 1e-300, 1e300, 2.2250738585072014e-308, 0.5, 2.0, 3.0})
        CHECK(ulp_diff(fm::log_t(xx, fm::bs_table_host() + 288), std::log(xx)) <= 1.0);
    CHECK(e_expt <= 1.0);
    CHECK(fm::exp_t(-800.0, fm::bs_table_host() + 256) == 0.0 && fm::exp_t(0.0, fm::bs_table_host() + 256) == 1.0);
    for (double xx : {-708.0, -707.9, -1e-300, 1e-17, 0.0108, 709.0, 700.5}) CHECK(ulp_diff(fm::exp_t(xx, fm::bs_table_host() + 256), std::exp(xx)) <= 1.0);
    CHECK(e_log <= 1.0);
    CHECK(e_div <= 1.0);
    CHECK(e_sqrt <= 1.0);
    CHECK(e_E <= 1.0);
    CHECK(e_phi_abs <= 4e-16);
    CHECK(e_phi_tail <= 8.0);   // far lower tail, relative; the reference's 0.5*(erf+1) is only ABSOLUTE-accurate there
    // edge cases
    CHECK(fm::exp_(-800.0) == 0.0);
    CHECK(fm::exp_(0.0) == 1.0);
    CHECK(fm::log_(1.0) == 0.0);
    double P, N, E;
    fm::phi_pair(0.0, P, N, E);
    CHECK(std::fabs(P - 0.5) <= 1.2e-16 && std::fabs(N - 0.5) <= 1.2e-16 && E == 1.0);
    fm::phi_pair(40.0, P, N, E);
    CHECK(P == 1.0 && N == 0.0 && E == 0.0);
    fm::phi_pair_pw(0.0, fm::bs_table_host(), P, N, E);
    CHECK(std::fabs(P - 0.5) <= 1.2e-16 && std::fabs(N - 0.5) <= 1.2e-16 && E == 1.0);
    fm::phi_pair_pw(40.0, fm::bs_table_host(), P, N, E);
    CHECK(P == 1.0 && N == 0.0 && E == 0.0);
    fm::phi_pair_pw(-1e-300, fm::bs_table_host(), P, N, E);
    CHECK(std::fabs(P - 0.5) <= 1.2e-16);
    return report("test_fastmath");
}
