// CPU-only: the expression TYPE layer (no CUDA call).  Mirrors the static checks of the
// reference's type traits (test/util/type_traits_unittest.cpp) and the host-side planning.
// NDEBUG: the reference guards shape mismatches with assert only (binary.hpp:86-90); with asserts
// compiled out the C ABI still rejects them, which is what is checked below.
#define NDEBUG
#include <fastad>

#include "check.hpp"

using namespace ad;

int main()
{
    // fake device addresses: nothing is dereferenced on the host
    double* base = reinterpret_cast<double*>(0x700000000000ull);
    VarView<double, vec> x(base, base + 4096, 1000), y(base + 8192, base + 12288, 1000);
    VarView<double> s(base + 20000, base + 20001);
    VarView<double, mat> M(base + 30000, base + 40000, 1000, 64);
    VarView<double, vec> w(base + 50000, base + 50100, 64);

    // shapes propagate like shape_traits.hpp:102-110
    auto e1 = ad::sin(x) * y + s;
    static_assert(std::is_same_v<decltype(e1)::shape_t, vec>);
    auto e2 = ad::sum(e1);
    static_assert(std::is_same_v<decltype(e2)::shape_t, scl>);
    auto e3 = ad::dot(M, w);
    static_assert(std::is_same_v<decltype(e3)::shape_t, vec>);
    CHECK(e3.rows() == 1000 && e3.cols() == 1);

    // eager folding of scalar constants (unary.hpp:182-184, binary.hpp:251-256, pow.hpp:219-227)
    auto c = ad::exp(ad::constant(2.0)) * 3.0 + ad::pow<2>(ad::constant(3.0));
    static_assert(core::is_scl_constant_v<decltype(c)>);
    CHECK_DOUBLE_EQ(c.feval(), std::exp(2.0) * 3.0 + 9.0);
    auto c2 = ad::sum(ad::constant(4.0));
    static_assert(core::is_scl_constant_v<decltype(c2)>);

    // arithmetic operands convert to Constant<scl> (type_traits.hpp:214-218)
    auto e4 = 0.5 * (ad::erf(x / std::sqrt(2.)) + 1.);
    static_assert(std::is_same_v<decltype(e4)::shape_t, vec>);

    // bind_cache_size: the fused plan materialises only the root (the reference: 5 N-sized caches)
    auto sz = ad::bind_cache_size(ad::sum(ad::exp(x) * y));
    CHECK(sz(0) == 1 && sz(1) == 0);
    // a reduction nested under scalar nodes costs one boundary scalar
    Var<double>* dummy = nullptr; (void)dummy;
    VarView<double> w4(base + 60000, base + 60001), w5(base + 60002, base + 60003);
    auto nested = (w4 = ad::sum(x * x), w5 = w4 * w4 + ad::sin(w4));
    auto sz2 = ad::bind_cache_size(nested);
    CHECK(sz2(0) == 2 && sz2(1) == 1);
    // normal(y, dot(X, w), sigma): dot output is the only N-sized boundary
    auto reg = ad::normal_adj_log_pdf(ad::constant_view(base + 70000, 1000), ad::dot(ad::constant_view(base + 80000, 1000, 64), w), s);
    auto sz3 = ad::bind_cache_size(reg);
    CHECK(sz3(0) == 1001 && sz3(1) == 1000);

    // shape errors surface as exceptions carrying the library message
    VarView<double, vec> z(base + 90000, base + 91000, 7);
    bool threw = false;
    try { (void)ad::bind_cache_size(x + z); } catch (const ad::Error& e) { threw = std::strstr(e.what(), "shapes differ") != nullptr; }
    CHECK(threw);

    // sub-views alias the parent storage (var_view.hpp:154-172)
    auto h = x.head(10), t = x.tail(10);
    CHECK(h.data() == x.data() && t.data() == x.data() + 990 && x(5).data() == x.data() + 5);
    return report("test_types");
}
