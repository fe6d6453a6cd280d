// GPU, compiled with nvcc: the type-fused path (include/fastad_b200/fused.cuh) against the C-ABI
// path (ad::bind -> interpreter / hand kernels) on the same expressions and inputs.
#include <fastad>
#include <fastad_b200/fused.cuh>
#include <random>

#include "check.hpp"

using namespace ad;

static std::mt19937_64 gen(20261019);
static void fill(Var<double, vec>& v, double lo, double hi)
{
    std::uniform_real_distribution<double> U(lo, hi);
    for (size_t i = 0; i < v.size(); ++i) v.get()(i) = U(gen);
}
template <class A, class B>
static void expect_same(const A& a, const B& b, double rel, const char* what)
{
    double worst = 0, scale = 0;
    for (size_t i = 0; i < a.size(); ++i) { scale = std::max(scale, std::fabs(b(i))); worst = std::max(worst, std::fabs(a(i) - b(i))); }
    ++g_checks;
    if (!(worst <= rel * std::max(scale, 1e-300))) { ++g_failures; std::printf("FAIL %s: max abs diff %.3g (scale %.3g)\n", what, worst, scale); }
}

template <class P> auto Phi(const P& x) { return 0.5 * (ad::erf(x / std::sqrt(2.)) + 1.); }

int main()
{
    if (fadb_init(0) != 0) { std::printf("no CUDA device: %s\n", fadb_last_error()); return 2; }
    const size_t n = 100003;

    // ---- sum of a 3-deep elementwise chain with two scalar variables (reduction channels)
    {
        Var<double, vec> x(n), x2(n), c(n);
        fill(x, 0.5, 2.0); fill(c, 0.5, 2.0);
        for (size_t i = 0; i < n; ++i) x2.get()(i) = x.get()(i);
        Var<double> w0(2.0), w1(1.3), u0(2.0), u1(1.3);
        auto cv = ad::constant_view(c.data(), n);
        c.upload();
        auto f = ad::fuse(ad::sum(-0.5 * ad::pow<2>((x - w0 * cv) / w1)));
        auto b = ad::bind(ad::sum(-0.5 * ad::pow<2>((x2 - u0 * cv) / u1)));
        const double vf = ad::autodiff(f, 0.75), vb = ad::autodiff(b, 0.75);
        CHECK_REL(vf, vb, 1e-12);
        expect_same(x.get_adj(), x2.get_adj(), 1e-14, "chain x_adj");
        CHECK_REL(w0.get_adj(), u0.get_adj(), 1e-11);
        CHECK_REL(w1.get_adj(), u1.get_adj(), 1e-11);
        // adjoints accumulate on a second sweep
        const double a0 = x.get_adj()(5);
        ad::autodiff(f, 0.75);
        CHECK_REL(x.get_adj()(5), 2 * a0, 1e-15);
    }
    // ---- batched Black-Scholes call expression (example/black_scholes.cpp) with vec leaves
    {
        const size_t m = 4099;
        auto build = [&](Var<double, vec>& S, Var<double, vec>& sg, Var<double, vec>& r, Var<double, vec>& K,
                         Var<double, vec>& tau, Var<double, vec>& sq, std::vector<Var<double, vec>>& cache) {
            auto Kc = ad::constant_view(K.data(), m), tc = ad::constant_view(tau.data(), m), sc = ad::constant_view(sq.data(), m);
            auto common = (cache[0] = ad::log(S / Kc),
                           cache[1] = (cache[0] + ((r + sg * sg / 2.) * tc)) / (sg * sc),
                           cache[2] = cache[1] - (sg * sc));
            auto PV = Kc * ad::exp(-r * tc);
            return (common, Phi(cache[1]) * S - Phi(cache[2]) * PV);
        };
        std::vector<Var<double, vec>> v1, v2;
        for (int k = 0; k < 2; ++k) {
            auto& v = k ? v2 : v1;
            for (int j = 0; j < 9; ++j) v.emplace_back(m);
        }
        fill(v1[0], 50, 150); fill(v1[1], 0.05, 0.6); fill(v1[2], 0.0, 0.08); fill(v1[3], 50, 150); fill(v1[4], 0.05, 2.0);
        for (size_t i = 0; i < m; ++i) v1[5].get()(i) = std::sqrt(v1[4].get()(i));
        for (int j = 0; j < 6; ++j) for (size_t i = 0; i < m; ++i) v2[j].get()(i) = v1[j].get()(i);
        for (auto* v : {&v1, &v2}) for (int j = 3; j < 6; ++j) (*v)[j].upload();
        std::vector<Var<double, vec>> c1{v1[6], v1[7], v1[8]}, c2{v2[6], v2[7], v2[8]};
        auto f = ad::fuse(build(v1[0], v1[1], v1[2], v1[3], v1[4], v1[5], c1));
        auto b = ad::bind(build(v2[0], v2[1], v2[2], v2[3], v2[4], v2[5], c2));
        Var<double, vec> ones(m);
        for (size_t i = 0; i < m; ++i) ones.get()(i) = 1.0;
        ones.upload();
        ad::autodiff(f, ad::device_seed{ones.data()});
        auto pb = ad::autodiff(b, std::vector<double>(m, 1.0));
        std::vector<double> pf(m);
        detail::check(fadb_d2h(pf.data(), f.value_device(), m * sizeof(double)));
        double worst = 0;
        for (size_t i = 0; i < m; ++i) worst = std::max(worst, std::fabs(pf[i] - pb[i]));
        CHECK(worst <= 1e-12 * 150);
        expect_same(v1[0].get_adj(), v2[0].get_adj(), 1e-12, "bs dS");
        expect_same(v1[1].get_adj(), v2[1].get_adj(), 1e-12, "bs dsigma");
        expect_same(v1[2].get_adj(), v2[2].get_adj(), 1e-12, "bs dr");
        expect_same(c1[1].get(), c2[1].get(), 1e-14, "bs cache[1] value");
    }
    // ---- normal log-pdf with an expression mean (univariate regression) and scalar sigma
    {
        Var<double, vec> y(n), xx(n);
        fill(y, -1, 1); fill(xx, 0.5, 2.0);
        y.upload(); xx.upload();
        auto yc = ad::constant_view(y.data(), n), xc = ad::constant_view(xx.data(), n);
        Var<double> a1(0.3), b1(-0.2), s1(1.7), a2(0.3), b2(-0.2), s2(1.7);
        auto f = ad::fuse(ad::normal_adj_log_pdf(yc, a1 + b1 * xc, s1));
        auto b = ad::bind(ad::normal_adj_log_pdf(yc, a2 + b2 * xc, s2));
        CHECK_REL(ad::autodiff(f), ad::autodiff(b), 1e-12);
        CHECK_REL(a1.get_adj(), a2.get_adj(), 1e-11);
        CHECK_REL(b1.get_adj(), b2.get_adj(), 1e-11);
        CHECK_REL(s1.get_adj(), s2.get_adj(), 1e-11);
        // sigma <= 0: value -inf, adjoints untouched (normal.hpp:147, 160)
        s1.get() = -1.0;
        const double before = a1.get_adj();
        CHECK(ad::autodiff(f) == -std::numeric_limits<double>::infinity());
        CHECK(a1.get_adj() == before);
    }
    // ---- normal vvv with a variable vector sigma: validity pass, then -inf / no adjoints when invalid
    {
        const size_t m = 5001;
        Var<double, vec> x(m), mu(m), sg(m), x2(m), mu2(m), sg2(m);
        fill(x, -1, 1); fill(mu, -1, 1); fill(sg, 0.5, 2.0);
        for (size_t i = 0; i < m; ++i) { x2.get()(i) = x.get()(i); mu2.get()(i) = mu.get()(i); sg2.get()(i) = sg.get()(i); }
        auto f = ad::fuse(ad::normal_adj_log_pdf(x, mu, sg));
        auto b = ad::bind(ad::normal_adj_log_pdf(x2, mu2, sg2));
        CHECK_REL(ad::autodiff(f), ad::autodiff(b), 1e-12);
        expect_same(x.get_adj(), x2.get_adj(), 0.0, "vvv x_adj (bit exact)");
        expect_same(sg.get_adj(), sg2.get_adj(), 2.3e-16, "vvv sigma_adj (1 ulp of the largest)");
        x.reset_adj(); mu.reset_adj(); sg.reset_adj();
        sg.get()(123) = 0.0;
        CHECK(ad::autodiff(f) == -std::numeric_limits<double>::infinity());
        CHECK(x.get_adj()(7) == 0.0 && sg.get_adj()(124) == 0.0);
    }
    // ---- scalar expression (n = 1), seed 1
    {
        Var<double> p(1.5928), q(-0.291), r(5.1023);
        auto f = ad::fuse(p * r + ad::sin(ad::cos(p + q)) * q - p / ad::exp(r));
        const double x1 = 1.5928, x2 = -0.291, x3 = 5.1023;
        CHECK_DOUBLE_EQ(ad::autodiff(f), x1 * x3 + std::sin(std::cos(x1 + x2)) * x2 - x1 / std::exp(x3));
        CHECK_DOUBLE_EQ(p.get_adj(), x3 - x2 * std::cos(std::cos(x1 + x2)) * std::sin(x1 + x2) - std::exp(-x3));
        CHECK_DOUBLE_EQ(r.get_adj(), x1 + x1 * std::exp(-x3));
    }
    return report("test_fused");
}
