// GPU: the ad:: host layer end to end, mirroring the reference's own tests
// (test/reverse/core/{eval,bind}_unittest.cpp, test/integration/{node,ad}_inttest.cpp,
// test/reverse/stat/normal_unittest.cpp) and the README worked examples.
#include <fastad>
#include <array>
#include <random>

#include "check.hpp"

using namespace ad;

static void eval_and_bind()
{
    // eval_unittest.cpp:32-48 with MockUnary f(x) = 2x
    Var<double> x(2.31);
    core::UnaryNode<core::MockUnary, VarView<double>> expr(x);
    auto b = ad::bind(expr);
    CHECK_DOUBLE_EQ(ad::autodiff(b), 4.62);
    CHECK_DOUBLE_EQ(x.get_adj(), 2.);
    ad::evaluate_adj(b);
    CHECK_DOUBLE_EQ(x.get_adj(), 4.);       // adjoints accumulate

    // bind_unittest.cpp:13-35
    Var<double> w1(1.), w2(2.), w3, w4;
    auto e = ad::bind((w3 = w1 * w2, w4 = w3 * w3));
    for (int rep = 0; rep < 2; ++rep) {
        CHECK_DOUBLE_EQ(ad::autodiff(e), 4.);
        CHECK_DOUBLE_EQ(w4.get_adj(), 1.);
        CHECK_DOUBLE_EQ(w3.get_adj(), 4.);
        CHECK_DOUBLE_EQ(w2.get_adj(), 4.);
        CHECK_DOUBLE_EQ(w1.get_adj(), 8.);
        CHECK_DOUBLE_EQ(w3.get(), 2.);
        CHECK_DOUBLE_EQ(w4.get(), 4.);
        w1.reset_adj(); w2.reset_adj(); w3.reset_adj(); w4.reset_adj();
    }
    // changing a value through the host mirror is seen by the next evaluation
    w1.get() = 3.;
    CHECK_DOUBLE_EQ(ad::autodiff(e), 36.);
}

static void node_integration()
{
    // node_inttest.cpp:27-159
    {
        Var<double> x(3.1);
        auto expr = ad::sin(x);
        CHECK_DOUBLE_EQ(ad::autodiff(expr), std::sin(3.1));
        CHECK_DOUBLE_EQ(x.get_adj(0, 0), std::cos(3.1));
    }
    {
        Var<double> x(3.1);
        auto expr = -x;
        CHECK_DOUBLE_EQ(ad::autodiff(expr), -3.1);
        CHECK_DOUBLE_EQ(x.get_adj(0, 0), -1.);
    }
    {
        Var<double> x(3.1);
        auto expr = ad::sin(ad::log(x));
        CHECK_DOUBLE_EQ(ad::autodiff(expr), std::sin(std::log(3.1)));
        CHECK_DOUBLE_EQ(x.get_adj(0, 0), std::cos(std::log(3.1)) / 3.1);
    }
    {
        Var<double> x(1.), y(2.), z(3.);
        auto expr = (x + y) - z;
        CHECK_DOUBLE_EQ(ad::autodiff(expr), 0.);
        CHECK_DOUBLE_EQ(x.get_adj(), 1.);
        CHECK_DOUBLE_EQ(y.get_adj(), 1.);
        CHECK_DOUBLE_EQ(z.get_adj(), -1.);
    }
    {
        double x1 = 1.2041, x2 = -2.2314;
        Var<double> leaf1(x1), leaf2(x2);
        auto expr = leaf1 * leaf2 + ad::sin(leaf1 + leaf2) * leaf2 - leaf1 / leaf2;
        auto b = ad::bind(expr);
        CHECK_DOUBLE_EQ(ad::evaluate(b), x1 * x2 + std::sin(x1 + x2) * x2 - x1 / x2);
        ad::evaluate_adj(b, 1);
        CHECK_DOUBLE_EQ(leaf1.get_adj(), x2 + std::cos(x1 + x2) * x2 - 1. / x2);
        CHECK_DOUBLE_EQ(leaf2.get_adj(), x1 + std::cos(x1 + x2) * x2 + std::sin(x1 + x2) + x1 / (x2 * x2));
    }
    {
        double x1 = 1.5928, x2 = -0.291, x3 = 5.1023;
        Var<double> leaf1(x1), leaf2(x2), leaf3(x3);
        auto expr = leaf1 * leaf3 + ad::sin(ad::cos(leaf1 + leaf2)) * leaf2 - leaf1 / ad::exp(leaf3);
        CHECK_DOUBLE_EQ(ad::autodiff(expr), x1 * x3 + std::sin(std::cos(x1 + x2)) * x2 - x1 / std::exp(x3));
        CHECK_DOUBLE_EQ(leaf1.get_adj(), x3 - x2 * std::cos(std::cos(x1 + x2)) * std::sin(x1 + x2) - std::exp(-x3));
        CHECK_DOUBLE_EQ(leaf2.get_adj(), std::sin(std::cos(x1 + x2)) - x2 * std::cos(std::cos(x1 + x2)) * std::sin(x1 + x2));
        CHECK_DOUBLE_EQ(leaf3.get_adj(), x1 + x1 * std::exp(-x3));
    }
    {   // node_inttest.cpp:380-400: sum(v * v) over a 3x2 matrix == 29 (1+4+1+9+4+9 = 28? data below)
        Var<double, mat> v(3, 2);
        v.get() = {1., -1., 2., 2., 3., -3.};
        auto expr = ad::sum(v * v);
        CHECK_DOUBLE_EQ(ad::autodiff(expr), 28.);
        for (size_t j = 0; j < 2; ++j)
            for (size_t i = 0; i < 3; ++i) CHECK_DOUBLE_EQ(v.get_adj(i, j), 2 * v.get(i, j));
    }
}

static void ad_integration()
{
    // ad_inttest.cpp:10-22, 66-140: F, G, H over x = [0.1, 2.3, -1, 4.1, -5.21]
    std::array<double, 5> val = {0.1, 2.3, -1., 4.1, -5.21};
    auto fresh = [&](std::vector<Var<double>>& v) {
        v.clear();
        for (double x : val) v.emplace_back(x);
    };
    std::vector<Var<double>> v;
    std::array<Var<double>, 10> w;
    {
        fresh(v);
        auto expr = (w[0] = ad::sin(v[0]) * ad::cos(v[1]), w[1] = v[2] + v[3] * v[4], w[2] = w[0] + w[1] + ad::constant(3.14));
        ad::autodiff(expr);
        CHECK_DOUBLE_EQ(v[0].get_adj(), std::cos(val[0]) * std::cos(val[1]));
        CHECK_DOUBLE_EQ(v[1].get_adj(), -std::sin(val[0]) * std::sin(val[1]));
        CHECK_DOUBLE_EQ(v[2].get_adj(), 1);
        CHECK_DOUBLE_EQ(v[3].get_adj(), val[4]);
        CHECK_DOUBLE_EQ(v[4].get_adj(), val[3]);
    }
    {
        fresh(v);
        for (auto& wi : w) wi.reset_adj();
        auto expr = (w[0] = ad::sum(v.begin(), v.end(), [](const auto& var) { return ad::sin(var); }),
                     w[1] = w[0] * w[0] - ad::sum(v.begin(), v.end(), [](const auto& var) { return ad::cos(var); }));
        ad::autodiff(expr);
        double sum = 0;
        for (double x : val) sum += std::sin(x);
        // `sum` is a cancelling sum of +-0.8 terms: 4 ulp of the TERMS, not of the result
        // (libdevice sin/cos differ from glibc's by <= 1 ulp)
        double abs_sum = 0;
        for (double x : val) abs_sum += std::fabs(std::sin(x));
        for (size_t j = 0; j < 5; ++j)
            CHECK_NEAR(v[j].get_adj(), 2 * sum * std::cos(val[j]) + std::sin(val[j]),
                       4 * 2.3e-16 * (2 * abs_sum * std::fabs(std::cos(val[j])) + std::fabs(std::sin(val[j]))));
    }
    {
        fresh(v);
        for (auto& wi : w) wi.reset_adj();
        auto expr = (w[0] = v[0] * v[4]);
        ad::autodiff(expr);
        CHECK_DOUBLE_EQ(v[0].get_adj(), val[4]);
        CHECK_DOUBLE_EQ(v[1].get_adj(), 0);
        CHECK_DOUBLE_EQ(v[4].get_adj(), val[0]);
    }
    {   // ad_inttest.cpp:143-156: 1000 scalar variables through sum/prod lambdas (crash test)
        constexpr size_t n = 1000;
        std::vector<Var<double>> x;
        std::default_random_engine gen;
        std::normal_distribution<double> dist(0., 1.);
        x.reserve(n);
        for (size_t i = 0; i < n; ++i) x.emplace_back(dist(gen));
        for (auto& wi : w) wi.reset_adj();
        auto expr = ad::bind((
            w[0] = ad::sin(x[0]) * ad::cos(ad::exp(x[1])) + ad::exp(x[0]) - x[1],
            w[1] = ad::sin(w[0]) - ad::sum(x.begin(), x.end(), [](const auto& var) { return ad::cos(var) * ad::exp(var); }),
            w[2] = ad::sin(w[1]) + ad::sum(x.begin(), x.end(), [](const auto& var) { return ad::sin(var) * ad::exp(var); }),
            w[3] = ad::sin(w[2]) + ad::prod(x.begin(), x.end(), [](const auto& var) { return ad::cos(var); })));
        double f = ad::autodiff(expr);
        CHECK(std::isfinite(f));
        // d/dx_k of the last sum term alone, through sin(w2): spot-check one adjoint by finite differences
        const size_t k = 17;
        const double g = x[k].get_adj(), x0 = x[k].get(), h = 1e-6;
        for (auto& xi : x) xi.reset_adj();
        for (auto& wi : w) wi.reset_adj();
        x[k].get() = x0 + h; double fp = ad::autodiff(expr);
        x[k].get() = x0 - h; double fm = ad::autodiff(expr);
        CHECK_NEAR(g, (fp - fm) / (2 * h), 1e-6 * std::max(1.0, std::fabs(g)));
    }
}

static void stat_normal()
{
    // normal_unittest.cpp:105-297 through Var-owned storage
    Var<double> sx(2.31), smu(-0.2), ssig(0.01);
    Var<double, vec> vx(3), vmu(3), vsig(3);
    vx.get() = {3.1, -2.3, 1.3};
    vmu.get() = {-0.3, -2.3, -1.2};
    vsig.get() = {0.01, 1.03, 2.41};
    {
        auto b = ad::bind(ad::normal_adj_log_pdf(sx, smu, ssig));
        CHECK_DOUBLE_EQ(ad::autodiff(b), -31495.8948298140203406);
        CHECK_DOUBLE_EQ(sx.get_adj(), -25100.);
        CHECK_DOUBLE_EQ(smu.get_adj(), 25100.);
        CHECK_DOUBLE_EQ(ssig.get_adj(), 6300000.0000000009313226);
        sx.reset_adj(); smu.reset_adj(); ssig.reset_adj();
    }
    {
        auto b = ad::bind(ad::normal_adj_log_pdf(vx, smu, ssig));
        CHECK_DOUBLE_EQ(ad::autodiff(b), -87736.1844894420355558);
        CHECK_DOUBLE_EQ(vx.get_adj(0, 0), -33000.);
        CHECK_DOUBLE_EQ(vx.get_adj(1, 0), 20999.9999999999963620);
        CHECK_DOUBLE_EQ(vx.get_adj(2, 0), -15000.);
        CHECK_DOUBLE_EQ(smu.get_adj(), 27000.0000000000036380);
        CHECK_DOUBLE_EQ(ssig.get_adj(), 17549700.);
        vx.reset_adj(); smu.reset_adj(); ssig.reset_adj();
    }
    {
        auto b = ad::bind(ad::normal_adj_log_pdf(vx, vmu, vsig));
        CHECK_DOUBLE_EQ(ad::autodiff(b), -57796.8420570641465019);
        CHECK_DOUBLE_EQ(vx.get_adj(0, 0), -34000.);
        CHECK_DOUBLE_EQ(vx.get_adj(2, 0), -0.4304333603071572);
        CHECK_DOUBLE_EQ(vmu.get_adj(2, 0), 0.4304333603071572);
        CHECK_DOUBLE_EQ(vsig.get_adj(0, 0), 11559900.);
        CHECK_DOUBLE_EQ(vsig.get_adj(1, 0), -0.9708737864077670);
        CHECK_DOUBLE_EQ(vsig.get_adj(2, 0), 0.0315698758372999);
        vx.reset_adj(); vmu.reset_adj(); vsig.reset_adj();
        // not positive definite => -inf and no adjoint update (normal_unittest.cpp:266-272)
        vsig.get()(0) = 0;
        CHECK(ad::autodiff(b) == -std::numeric_limits<double>::infinity());
        CHECK(vx.get_adj(0, 0) == 0. && vsig.get_adj(1, 0) == 0.);
    }
}

static void readme_examples()
{
    // README.md:582-607 quadratic form x^T Sigma x (transpose node is out of scope: x^T is a 1x2 Var)
    {
        Var<double, mat> x(2, 1), xt(1, 2);
        x.get() = {0.5, 0.6};
        xt.get() = {0.5, 0.6};
        auto Sigma = ad::constant({2., 3., 3., 6.}, 2, 2);
        auto b = ad::bind(ad::dot(ad::dot(xt, Sigma), x));
        auto f = ad::autodiff(b, std::vector<double>{1.});
        CHECK_DOUBLE_EQ(f[0], 4.46);
        CHECK_DOUBLE_EQ(x.get_adj(0, 0) + xt.get_adj(0, 0), 5.6);
        CHECK_DOUBLE_EQ(x.get_adj(1, 0) + xt.get_adj(0, 1), 10.2);
    }
    // README.md:623-662 linear regression, row loop: loss 6655, adj [-1210, -12100]
    {
        const double X[5][2] = {{1, 10}, {2, 20}, {3, 30}, {4, 40}, {5, 50}};
        const double y[5] = {32, 64, 96, 128, 160};
        Var<double, mat> theta(2, 1);
        theta.get() = {1., 2.};
        Var<double, mat> xrow(1, 2), yrow(1, 1);      // device-resident row buffers
        auto xi = ad::constant_view(xrow.data(), 1, 2);
        auto yi = ad::constant_view(yrow.data(), 1, 1);
        auto expr = ad::bind(ad::pow<2>(yi - ad::dot(xi, theta)));
        double loss = 0;
        for (int i = 0; i < 5; ++i) {
            xrow.get() = {X[i][0], X[i][1]};
            yrow.get() = {y[i]};
            xrow.upload(); yrow.upload();
            loss += ad::autodiff(expr, std::vector<double>{1.})[0];
        }
        CHECK_DOUBLE_EQ(loss, 6655.);
        CHECK_DOUBLE_EQ(theta.get_adj(0, 0), -1210.);
        CHECK_DOUBLE_EQ(theta.get_adj(1, 0), -12100.);
    }
    // README.md:676-745 3-layer network, row loop; goldens test/util/GenTestData.nb:691-700
    {
        const double parm[25] = {
            0.04398400506863087, 0.9601256114996133, -0.5209410945118056, -0.8005256013595985,
            -0.028791364112533024, 0.6358093220436887, 0.5846033162880331, -0.4433822901819533,
            0.2243037009072486, 0.9750495768482685, -0.8240837134824699, 0.23629950009691214,
            0.6663923726986263, -0.4988280521070929, -0.7814276530077855, -0.9110531473087247,
            -0.23015593969194725, -0.13636652333316235, 0.26342534615305047, 0.8415352118678241,
            0.9203421214171716, 0.6562896796607487, 0.848247693352548, -0.7486969470512386,
            0.21522047772523667};
        Var<double, vec> P(25);
        for (int i = 0; i < 25; ++i) P.get()(i) = parm[i];
        P.upload();
        auto view = [&](size_t off, size_t r, size_t c) { return VarView<double, mat>(P.data() + off, P.data_adj() + off, r, c); };
        auto A1 = view(0, 3, 2), b1 = view(6, 3, 1), A2 = view(9, 3, 3), b2 = view(18, 3, 1), A3 = view(21, 1, 3), b3 = view(24, 1, 1);
        Var<double, mat> xrow(2, 1), yrow(1, 1);
        auto xi = ad::constant_view(xrow.data(), 2, 1);
        auto yi = ad::constant_view(yrow.data(), 1, 1);
        auto x1 = ad::dot(A1, xi) + b1;
        auto y1 = ad::sigmoid(x1);
        auto y2 = ad::sigmoid(ad::dot(A2, y1) + b2) + x1;
        auto y3 = ad::dot(A3, y2) + b3;
        auto expr = ad::bind(ad::pow<2>(yi - y3));
        double loss = 0;
        for (int i = 1; i <= 5; ++i) {
            xrow.get() = {1. * i, 10. * i};
            yrow.get() = {32. * i};
            xrow.upload(); yrow.upload();
            loss += ad::autodiff(expr, std::vector<double>{1.})[0];
        }
        P.invalidate_host();      // the adjoints were written through views of P's device buffers
        CHECK_REL(loss, 92030.04305391687, 1e-12);
        CHECK_REL(P.get_adj()(0), -2953.051309234279, 1e-12);
        CHECK_REL(P.get_adj()(3), -29530.51309234279, 1e-12);
    }
}

static void vectors_and_seeds()
{
    // batched expression over vec Vars with a vector seed (eval.hpp:68-73); head/tail sub-views
    const size_t n = 1000;
    Var<double, vec> x(n), y(n);
    for (size_t i = 0; i < n; ++i) { x.get()(i) = 0.5 + 1e-3 * i; y.get()(i) = 2.0 - 1e-3 * i; }
    auto b = ad::bind(ad::exp(x) * y);
    auto f = ad::autodiff(b, std::vector<double>(n, 2.0));
    CHECK_DOUBLE_EQ(f[10], std::exp(x.get()(10)) * y.get()(10));
    CHECK_DOUBLE_EQ(x.get_adj()(10), 2.0 * std::exp(x.get()(10)) * y.get()(10));
    CHECK_DOUBLE_EQ(y.get_adj()(999), 2.0 * std::exp(x.get()(999)));
    x.reset_adj(); y.reset_adj();
    auto s = ad::bind(ad::sum(x.head(10) * x.tail(10)));
    double v = ad::autodiff(s);
    double ref = 0;
    for (size_t i = 0; i < 10; ++i) ref += x.get()(i) * x.get()(n - 10 + i);
    CHECK_REL(v, ref, 1e-14);
    CHECK_DOUBLE_EQ(x.get_adj()(0), x.get()(n - 10));
    CHECK_DOUBLE_EQ(x.get_adj()(n - 1), x.get()(9));
    CHECK(x.get_adj()(500) == 0.);
    // ad::prod with an exact zero (prod.hpp:197-207; fixture vector base_fixture.hpp:109-115)
    Var<double, vec> p(5);
    p.get() = {3.1, -2.3, 1.3, 0., 5.1};
    CHECK(ad::autodiff(ad::prod(p), 2.0) == 0.);
    CHECK_REL(p.get_adj()(3), 2.0 * 3.1 * -2.3 * 1.3 * 5.1, 1e-15);
    CHECK(p.get_adj()(0) == 0.);
    // Var copies are deep (var.hpp:45-55); VarView copies alias
    Var<double> a(1.5);
    Var<double> c = a;
    c.get() = 2.5;
    CHECK(a.get() == 1.5);
    VarView<double> av = a;
    ad::autodiff(av * av);
    CHECK_DOUBLE_EQ(a.get_adj(), 3.0);
}

static void next_rows()
{
    // ---- ad::norm (norm_unittest.cpp:31-49) and a norm of an expression (node_inttest.cpp:402-442 shape)
    {
        Var<double, vec> v(5);
        v.get() = {3.1, -2.3, 1.3, 0., 5.1};
        core::UnaryNode<core::MockUnary, VarView<double, vec>> u(v);
        auto b = ad::bind(ad::norm(u));
        double ref = 0;
        for (double x : {3.1, -2.3, 1.3, 0., 5.1}) ref += 4 * x * x;
        CHECK_DOUBLE_EQ(ad::autodiff(b, 9.2313), ref);
        CHECK_DOUBLE_EQ(v.get_adj()(0), 2 * (9.2313 * 2 * (2 * 3.1)));
        CHECK_DOUBLE_EQ(v.get_adj()(3), 0.);
    }
    // ---- README.md:582-607 exactly as written there: x^T Sigma x with ad::transpose
    {
        Var<double, mat> x(2, 1);
        x.get() = {0.5, 0.6};
        auto Sigma = ad::constant({2., 3., 3., 6.}, 2, 2);
        auto b = ad::bind(ad::dot(ad::dot(ad::transpose(x), Sigma), x));
        auto f = ad::autodiff(b, std::vector<double>{1.});
        CHECK_DOUBLE_EQ(f[0], 4.46);
        CHECK_DOUBLE_EQ(x.get_adj(0, 0), 5.6);
        CHECK_DOUBLE_EQ(x.get_adj(1, 0), 10.2);
    }
    // ---- ad::if_else with comparison / logical nodes (if_else_unittest.cpp, binary.hpp:347-470)
    for (double a0 : {1.5, -0.5}) {
        Var<double> a(a0), b(0.25);
        Var<double, vec> x(3);
        x.get() = {0.5, 1.5, 2.5};
        auto e = ad::bind(ad::sum(ad::if_else((b < a) && (a != b), x * a, ad::exp(x * b))));
        const double v = ad::autodiff(e);
        if (a0 > 0.25) {
            CHECK_DOUBLE_EQ(v, 4.5 * a0);
            CHECK_DOUBLE_EQ(a.get_adj(), 4.5);
            CHECK_DOUBLE_EQ(b.get_adj(), 0.);
            CHECK_DOUBLE_EQ(x.get_adj()(1), a0);
        } else {
            CHECK_DOUBLE_EQ(v, std::exp(0.125) + std::exp(0.375) + std::exp(0.625));
            CHECK_DOUBLE_EQ(a.get_adj(), 0.);
            CHECK_DOUBLE_EQ(x.get_adj()(2), 0.25 * std::exp(0.625));
        }
    }
    // ---- op= nodes (eq_unittest.cpp:143-240)
    {
        const double seed = 3.2, x0 = 2.31;
        Var<double> w(x0);
        auto e1 = ad::bind((w *= w));
        CHECK_DOUBLE_EQ(ad::autodiff(e1, seed), x0 * x0);
        CHECK_DOUBLE_EQ(w.get_adj(), seed * 2 * x0);
        CHECK_DOUBLE_EQ(w.get(), x0);                       // the previous value is restored (eq.hpp:255)
        Var<double, vec> v(5);
        Var<double> s(x0);
        v.get() = {3.1, -2.3, 1.3, 0., 5.1};
        auto e2 = ad::bind((v *= s));
        auto r = ad::autodiff(e2, std::vector<double>{2.1, -0.3, 1.0, 0.5, 4.0});
        CHECK_DOUBLE_EQ(r[0], 3.1 * x0);
        CHECK_DOUBLE_EQ(v.get_adj()(4), 4.0 * x0);
        CHECK_DOUBLE_EQ(s.get_adj(), 2.1 * 3.1 + 0.3 * 2.3 + 1.0 * 1.3 + 0.5 * 0. + 4.0 * 5.1);
    }
    // ---- other diagonal log-pdfs (cauchy_unittest.cpp:80-262, uniform_unittest.cpp:80-104, bernoulli_unittest.cpp:148-153)
    {
        Var<double, vec> x(3), loc(3), sc(3);
        x.get() = {0.5, -1.3, -3.2414999};
        loc.get() = {0.4, -2.30000001, -10.32};
        sc.get() = {0.51, 0.01, 3.4};
        auto b = ad::bind(ad::cauchy_adj_log_pdf(x, loc, sc));
        CHECK_DOUBLE_EQ(ad::autodiff(b), -6.867595467569049816);
        CHECK_DOUBLE_EQ(x.get_adj()(1), -1.999800000003999267);
        CHECK_DOUBLE_EQ(loc.get_adj()(0), 0.740466493891151267);
        CHECK_DOUBLE_EQ(sc.get_adj()(1), 99.980002000199945655);
        Var<double> ux(0.45), lo(-3.2415), hi(0.5231);
        auto u = ad::bind(ad::uniform_adj_log_pdf(ux, lo, hi));
        CHECK_DOUBLE_EQ(ad::autodiff(u), -1.3256416139079406);
        CHECK_DOUBLE_EQ(lo.get_adj(), 1. / (0.5231 + 3.2415));
        ux.get() = -100;                                    // out of range: -inf, adjoints untouched
        CHECK(ad::autodiff(u) == -std::numeric_limits<double>::infinity());
        CHECK_DOUBLE_EQ(hi.get_adj(), -1. / (0.5231 + 3.2415));
        Var<double> p(0.0001);
        auto be = ad::bind(ad::bernoulli_adj_log_pdf(ad::constant(std::vector<double>{1, 0, 1}), p));
        CHECK_DOUBLE_EQ(ad::autodiff(be), -18.42078074895269779176);
        CHECK_DOUBLE_EQ(p.get_adj(), (2 - 3 * 0.0001) / (0.0001 * (1 - 0.0001)));
    }
    // ---- dense-covariance normal, vvm (normal_unittest.cpp:340-395); Sigma's lower triangle only
    {
        Var<double, vec> x(3), mu(3);
        Var<double, mat> Sig(3, 3);
        x.get() = {3.1, -2.3, 1.3};
        mu.get() = {-0.3, -2.3, -1.2};
        Sig.get() = {1.0, 0.3, 0.2, 0.0, 2.0, -0.3, 0.0, 0.0, 3.0};      // column-major, upper left at zero
        auto b = ad::bind(ad::normal_adj_log_pdf(x, mu, Sig));
        CHECK_DOUBLE_EQ(ad::autodiff(b), -7.3649692930088602);
        CHECK_DOUBLE_EQ(x.get_adj()(0), -3.4158218682114407);
        CHECK_DOUBLE_EQ(mu.get_adj()(1), -0.4279507603186097);
        CHECK_DOUBLE_EQ(Sig.get_adj(0, 0), 5.2989810672774862);
        CHECK_DOUBLE_EQ(Sig.get_adj(2, 1), -0.1530140218890801);
        CHECK_DOUBLE_EQ(Sig.get_adj(1, 2), -0.1530140218890801);
    }
    // ---- ad::det / ad::log_det (det_unittest.cpp:57-155, log_det_unittest.cpp:57-152)
    {
        const double seed = 9.2313;
        Var<double, mat> A(4, 4), B(4, 4);
        A.get() = {2, 3, -2, -1, 3, 5, 3, -1, 1, -1, 1, 2, 5, 3, 0, 7};          // init_fplu, column-major
        B.get() = {8, 3, 1, 5, 3, 5, 1, 3, 1, 1, 1, 0, 5, 3, 0, 7};              // init_llt (symmetric)
        auto d0 = ad::bind(ad::det(A));
        CHECK_DOUBLE_EQ(ad::autodiff(d0, seed), 153.);
        // adj = seed * det * inverse().transpose(): cofactor(0,0) = det [[5,-1,3],[3,1,0],[-1,2,7]] = 77
        CHECK_NEAR(A.get_adj(0, 0), seed * 77., 3e-13 * 1000);
        auto d1 = ad::bind(ad::det<ad::DetLLT>(B));
        CHECK_DOUBLE_EQ(ad::autodiff(d1, seed), 65.);
        auto d2 = ad::bind(ad::det<ad::DetLDLT>(B));
        CHECK_DOUBLE_EQ(ad::autodiff(d2, seed), 65.);
        B.reset_adj();
        auto l1 = ad::bind(ad::log_det<ad::LogDetLLT>(B));
        CHECK_DOUBLE_EQ(ad::autodiff(l1, seed), std::log(65.));
        // adj = seed * inverse().transpose(); inv(B)(2,2) = cofactor / det, cofactor(2,2) = det [[8,3,5],[3,5,3],[5,3,7]] = 110
        CHECK_NEAR(B.get_adj(2, 2), seed * 110. / 65., 3e-13);
        auto l0 = ad::bind(ad::log_det(A));
        CHECK_DOUBLE_EQ(ad::autodiff(l0, 0.), std::log(153.));
        static_assert(std::is_same_v<decltype(ad::det(A))::shape_t, ad::scl>);
    }
    // ---- ad::wishart_adj_log_pdf (wishart_unittest.cpp:62-106)
    {
        Var<double, mat> X(3, 3), V(3, 3);
        X.get() = {10, 2, 3, 2, 10, 1, 3, 1, 10};
        V.get() = {5, 1, 0, 1, 5, 1, 0, 1, 5};
        size_t n = 4;
        auto w = ad::bind(ad::wishart_adj_log_pdf(X, V, n));
        CHECK_DOUBLE_EQ(ad::autodiff(w), -12.55942947411780252764);
        // dX = 0.5 ((n-p-1) X^-1 - V^-1) = -0.5 V^-1 for n = 4, p = 3; V^-1(0,0) = 24/115
        CHECK_NEAR(X.get_adj(0, 0), -0.5 * 24. / 115., 1e-15);
        V.get() = {0, 0, 0, 0, 0, 0, 0, 0, 0};                                  // feval_invalid
        CHECK(ad::autodiff(w) == -std::numeric_limits<double>::infinity());
    }
    // ---- ad::for_each (for_each.hpp:19-131): evaluate all, value of the last, seed to the last only
    {
        std::vector<Var<double>> xs, ws(3);
        for (double x : {0.5, 1.5, 2.5}) xs.emplace_back(x);
        std::vector<size_t> idx{0, 1, 2};
        auto e = ad::bind(ad::for_each(idx.begin(), idx.end(), [&](size_t k) { return ws[k] = xs[k] * xs[k]; }));
        CHECK_DOUBLE_EQ(ad::autodiff(e), 6.25);
        CHECK_DOUBLE_EQ(ws[0].get(), 0.25);
        CHECK_DOUBLE_EQ(xs[2].get_adj(), 5.0);
        CHECK_DOUBLE_EQ(xs[0].get_adj(), 0.);
    }
}

// The README regression in its batched form and a vector model with many scalar parameters: both are
// bound from plain g++ code and must run on the one-pass least-squares kernel / a specialised stage kernel.
static void batched_models()
{
    // ---- README.md:623-662 with the 5 rows replicated 2000 times: one autodiff instead of a row loop
    {
        const size_t reps = 2000, rows = 5 * reps;
        const double Xr[5][2] = {{1, 10}, {2, 20}, {3, 30}, {4, 40}, {5, 50}};
        const double yr[5] = {32, 64, 96, 128, 160};
        Var<double, mat> X(rows, 2);
        Var<double, vec> y(rows);
        for (size_t i = 0; i < rows; ++i) {
            X.get()(i, 0) = Xr[i % 5][0]; X.get()(i, 1) = Xr[i % 5][1];
            y.get()(i) = yr[i % 5];
        }
        X.upload(); y.upload();
        Var<double, vec> theta(2);
        theta.get() = {1., 2.};
        auto expr = ad::bind(ad::sum(ad::pow<2>(ad::constant_view(y.data(), rows) - ad::dot(ad::constant_view(X.data(), rows, 2), theta))));
        unsigned long long j0[4], j1[4];
        fadb_jit_stats(j0);
        const unsigned long long l0 = fadb_launch_count();
        CHECK_REL(ad::autodiff(expr), 6655. * reps, 1e-13);
        CHECK(fadb_launch_count() - l0 == 2);                 // fadb_lsq: one pass over X + the scalar epilogue
        CHECK_REL(theta.get_adj()(0), -1210. * reps, 1e-13);
        CHECK_REL(theta.get_adj()(1), -12100. * reps, 1e-13);
        fadb_jit_stats(j1);
        CHECK(j1[3] == j0[3]);                                // no stage kernel involved
    }
    // ---- GLM shapes written with the reference's API take the one-pass route (fast=6): logistic regression with an intercept,
    //      sum(y * eta - log(1 + exp(eta))), eta = dot(X, beta) + b, and normal(y, dot(X, beta) + b, sigma)
    {
        const size_t rows = 4096, cols = 3;
        Var<double, mat> X(rows, cols);
        Var<double, vec> y(rows), yn(rows);
        std::vector<double> Xh(rows * cols), yh(rows), ynh(rows);
        unsigned long long st = 12345;
        auto rnd = [&]() { st = st * 6364136223846793005ull + 1442695040888963407ull; return double(st >> 11) / double(1ull << 53); };
        for (size_t j = 0; j < cols; ++j)
            for (size_t i = 0; i < rows; ++i) { Xh[i + j * rows] = rnd() - 0.5; X.get()(i, j) = Xh[i + j * rows]; }
        for (size_t i = 0; i < rows; ++i) { yh[i] = rnd() < 0.5 ? 0.0 : 1.0; ynh[i] = 2.0 * rnd(); y.get()(i) = yh[i]; yn.get()(i) = ynh[i]; }
        X.upload(); y.upload(); yn.upload();
        Var<double, vec> beta(cols);
        beta.get() = {0.3, -0.2, 0.1};
        Var<double> b(0.25), sigma(1.5);
        auto Xc = ad::constant_view(X.data(), rows, cols);
        auto eta = ad::dot(Xc, beta) + b;
        auto ll = ad::bind(ad::sum(ad::constant_view(y.data(), rows) * eta - ad::log(1. + ad::exp(eta))));
        unsigned long long j0[4], j1[4];
        fadb_jit_stats(j0);
        const unsigned long long l0 = fadb_launch_count();
        const double v = ad::autodiff(ll);
        fadb_jit_stats(j1);
        CHECK(fadb_launch_count() - l0 == 2 && j1[3] == j0[3] + 1 && j1[2] == j0[2]);    // the tile kernel + its scalar epilogue
        double ref = 0, gb = 0, g[3] = {0, 0, 0};
        for (size_t i = 0; i < rows; ++i) {
            double e = 0.25;
            for (size_t j = 0; j < cols; ++j) e += Xh[i + j * rows] * (j == 0 ? 0.3 : j == 1 ? -0.2 : 0.1);
            ref += yh[i] * e - std::log(1. + std::exp(e));
            const double d = yh[i] - 1. / (1. + std::exp(-e));
            gb += d;
            for (size_t j = 0; j < cols; ++j) g[j] += d * Xh[i + j * rows];
        }
        CHECK_REL(v, ref, 1e-12);
        CHECK_REL(b.get_adj(), gb, 1e-10);
        for (size_t j = 0; j < cols; ++j) CHECK_REL(beta.get_adj()(j), g[j], 1e-10);

        beta.reset_adj(); b.reset_adj();
        auto nl = ad::bind(ad::normal_adj_log_pdf(ad::constant_view(yn.data(), rows), ad::dot(Xc, beta) + b, sigma));
        fadb_jit_stats(j0);
        const double vn = ad::autodiff(nl);
        fadb_jit_stats(j1);
        CHECK(j1[3] == j0[3] + 1 && j1[2] == j0[2]);
        double ss = 0, gs_b = 0;
        for (size_t i = 0; i < rows; ++i) {
            double m = 0.25;
            for (size_t j = 0; j < cols; ++j) m += Xh[i + j * rows] * (j == 0 ? 0.3 : j == 1 ? -0.2 : 0.1);
            ss += (ynh[i] - m) * (ynh[i] - m);
            gs_b += (ynh[i] - m) / (1.5 * 1.5);
        }
        CHECK_REL(vn, -0.5 * ss / (1.5 * 1.5) - double(rows) * std::log(1.5), 1e-12);
        CHECK_REL(b.get_adj(), gs_b, 1e-10);
        CHECK_REL(sigma.get_adj(), (ss / (1.5 * 1.5) - double(rows)) / 1.5, 1e-10);
    }
    // ---- 12 scalar parameters broadcast into one vector stage (more reduction channels than the interpreter has)
    {
        const size_t n = 4321;
        Var<double, vec> x(n);
        for (size_t i = 0; i < n; ++i) x.get()(i) = 0.2 + 1.3 * double(i) / n;
        std::vector<Var<double>> p;
        for (int k = 0; k < 12; ++k) p.emplace_back(0.5 + 0.1 * k);
        auto term = [&](int k) { return p[k] * ad::sin((0.3 * (k + 1)) * x); };
        auto model = ad::sum(term(0) + term(1) + term(2) + term(3) + term(4) + term(5) + term(6) + term(7) + term(8) + term(9) +
                             term(10) + term(11));
        auto expr = ad::bind(model);
        unsigned long long j0[4], j1[4];
        fadb_jit_stats(j0);
        const double v = ad::autodiff(expr);
        fadb_jit_stats(j1);
        CHECK(j1[3] == j0[3] + 1 && j1[2] == j0[2]);          // ran on a specialised kernel
        double ref = 0, dp3 = 0, dx7 = 0;
        for (size_t i = 0; i < n; ++i) {
            const double xi = 0.2 + 1.3 * double(i) / n;
            for (int k = 0; k < 12; ++k) ref += (0.5 + 0.1 * k) * std::sin(0.3 * (k + 1) * xi);
            dp3 += std::sin(0.3 * 4 * xi);
            if (i == 7) for (int k = 0; k < 12; ++k) dx7 += (0.5 + 0.1 * k) * 0.3 * (k + 1) * std::cos(0.3 * (k + 1) * xi);
        }
        CHECK_REL(v, ref, 1e-12);
        CHECK_REL(p[3].get_adj(), dp3, 1e-12);
        CHECK_REL(x.get_adj()(7), dx7, 1e-13);
    }
}

static void node_level_protocol()
{
    // README.md:284-306: manual bind_cache, then the node's own feval() / beval(seed)
    // (unary.hpp:63-89, binary.hpp:100-136) -- `sin(x) + cos(v)` is a vector expression
    Var<double> x(0.7);
    Var<double, vec> v(3);
    v.get() = {0.1, -0.4, 1.2};
    auto expr = (ad::sin(x) + ad::cos(v));
    auto size_pack = expr.bind_cache_size();
    std::vector<double> val_buf(size_pack(0)), adj_buf(size_pack(1));
    expr.bind_cache({val_buf.data(), adj_buf.data()});
    auto f = expr.feval();
    CHECK(f.size() == 3);
    for (int i = 0; i < 3; ++i) CHECK_DOUBLE_EQ(f[i], std::sin(0.7) + std::cos(v.get()(i)));
    CHECK(x.get_adj() == 0.);                                  // feval touches no adjoint
    expr.beval(std::vector<double>{1., 2., -1.});
    CHECK_DOUBLE_EQ(x.get_adj(), 2. * std::cos(0.7));          // seeds sum into the scalar (binary.hpp:266-338)
    for (int i = 0; i < 3; ++i) CHECK_DOUBLE_EQ(v.get_adj()(i), (i == 0 ? 1. : i == 1 ? 2. : -1.) * -std::sin(v.get()(i)));
    expr.beval(std::vector<double>{1., 2., -1.});              // adjoints accumulate (var_view.hpp:91-94)
    CHECK_DOUBLE_EQ(x.get_adj(), 4. * std::cos(0.7));
    // scalar expression: default seed 1; a copy of the node shares the binding; a changed value is seen
    Var<double> a(1.5), b(-0.5);
    auto e2 = a * b + ad::exp(a);
    CHECK_DOUBLE_EQ(e2.feval(), 1.5 * -0.5 + std::exp(1.5));
    auto e3 = e2;
    e3.beval();
    CHECK_DOUBLE_EQ(a.get_adj(), -0.5 + std::exp(1.5));
    CHECK_DOUBLE_EQ(b.get_adj(), 1.5);
    a.get() = 2.0;
    CHECK_DOUBLE_EQ(e2.feval(), 2.0 * -0.5 + std::exp(2.0));
    // op= statement: feval applies it, beval restores the previous value (eq.hpp:240-271)
    Var<double> w(3.0), y(1.25);
    auto e4 = (w *= y);
    CHECK_DOUBLE_EQ(e4.feval(), 3.75);
    CHECK_DOUBLE_EQ(w.get(), 3.75);
    e4.beval(2.0);
    CHECK_DOUBLE_EQ(w.get(), 3.0);
    CHECK_DOUBLE_EQ(y.get_adj(), 2.0 * 3.0);
}

int main()
{
    if (fadb_init(0) != 0) { std::printf("no CUDA device: %s\n", fadb_last_error()); return 2; }
    eval_and_bind();
    node_integration();
    ad_integration();
    stat_normal();
    readme_examples();
    vectors_and_seeds();
    next_rows();
    batched_models();
    node_level_protocol();
    return report("test_host_layer");
}
