// The usage patterns of the reference's example/main.cpp (forward mode; reverse mode with the manual
// bind_cache_size / bind_cache protocol and a raw autodiff; placeholders held in std::vector), with the same
// function f(x, y) = exp((x sin y + x y) x sin y) at the same point, results CHECKED against the closed form
// instead of printed.  The reference file itself also compiles unchanged against <fastad>.
#include <fastad>
#include <vector>

#include "check.hpp"

static const double X1 = -0.201, X2 = 1.2241;

static void closed_form(double& f, double& dx, double& dy)
{
    const double s = std::sin(X2), c = std::cos(X2);
    const double w3 = X1 * s, w4 = w3 + X1 * X2;
    f = std::exp(w4 * w3);
    // d(w4 w3)/dx = (s + y) x s + (x s + x y) s ;  d/dy = (x c + x) x s + (x s + x y) x c
    dx = f * ((s + X2) * w3 + w4 * s);
    dy = f * ((X1 * c + X1) * w3 + w4 * X1 * c);
}

static void forward()
{
    using namespace ad;
    ForwardVar<double> w1(X1), w2(X2);
    w1.get_adjoint() = 1;                       // direction e_x
    ForwardVar<double> w3 = w1 * sin(w2);
    auto w4 = w3 + w1 * w2;
    auto w5 = exp(w4 * w3);
    double f, dx, dy;
    closed_form(f, dx, dy);
    CHECK_REL(w5.get_value(), f, 1e-15);
    CHECK_REL(w5.get_adjoint(), dx, 1e-14);
}

static void reverse_simple()
{
    using namespace ad;
    Var<double> w1(X1), w2(X2), w3, w4, w5;
    auto expr = (w3 = w1 * sin(w2), w4 = w3 + w1 * w2, w5 = exp(w4 * w3));
    auto size_pack = expr.bind_cache_size();
    std::vector<double> val_buf(size_pack(0)), adj_buf(size_pack(1));
    auto end = expr.bind_cache({val_buf.data(), adj_buf.data()});
    CHECK(end.val == val_buf.data() + size_pack(0) && end.adj == adj_buf.data() + size_pack(1));
    const double v = autodiff(expr);
    double f, dx, dy;
    closed_form(f, dx, dy);
    CHECK_REL(v, f, 1e-14);
    CHECK_REL(w5.get(), f, 1e-14);
    CHECK_REL(w1.get_adj(0, 0), dx, 1e-13);
    CHECK_REL(w2.get_adj(0, 0), dy, 1e-13);
}

static void reverse_vec()
{
    using namespace ad;
    std::vector<Var<double>> x(0), w(3);
    const double x_val[] = {X1, X2};
    for (size_t i = 0; i < 2; ++i) x.emplace_back(x_val[i]);
    auto expr = (w[0] = x[0] * sin(x[1]), w[1] = w[0] + x[0] * x[1], w[2] = exp(w[1] * w[0]));
    auto size_pack = expr.bind_cache_size();
    std::vector<double> val_buf(size_pack(0)), adj_buf(size_pack(1));
    expr.bind_cache({val_buf.data(), adj_buf.data()});
    autodiff(expr);
    double f, dx, dy;
    closed_form(f, dx, dy);
    CHECK_REL(x[0].get_adj(0, 0), dx, 1e-13);
    CHECK_REL(x[1].get_adj(0, 0), dy, 1e-13);
}

int main()
{
    forward();
    if (fadb_init(0) != 0) { std::printf("no CUDA device: %s\n", fadb_last_error()); return 2; }
    reverse_simple();
    reverse_vec();
    return report("example_main");
}
