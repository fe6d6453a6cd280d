// CPU, g++: forward-mode ad::ForwardVar / core::DualNum against the reference's own unit tests
// (test/forward/core/forward_unittest.cpp:16-149, dualnum_unittest.cpp:18-41), same inputs and
// expected values, EXPECT_DOUBLE_EQ -> CHECK_DOUBLE_EQ (4 ulp).
#include <fastad>
#include <type_traits>

#include "check.hpp"

using ad::ForwardVar;

int main()
{
    // ---- dualnum_unittest.cpp
    {
        static_assert(std::is_same<ad::core::DualNum<double>::value_type, double>::value, "value_type");
        ad::core::DualNum<double> dual(2.1, 2.3);
        CHECK_DOUBLE_EQ(dual.get_value(), 2.1);
        CHECK_DOUBLE_EQ(dual.get_adjoint(), 2.3);
        dual.set_value(3.4);
        CHECK_DOUBLE_EQ(dual.get_value(), 3.4);
        CHECK_DOUBLE_EQ(dual.get_adjoint(), 2.3);
        dual.set_adjoint(3.4);
        CHECK_DOUBLE_EQ(dual.get_value(), 3.4);
        CHECK_DOUBLE_EQ(dual.get_adjoint(), 3.4);
    }
    // ---- forward_unittest.cpp, unary
    auto seeded = [](double v) { ForwardVar<double> x(v); x.set_adjoint(1); return x; };
    { ForwardVar<double> y = -seeded(2); CHECK_DOUBLE_EQ(y.get_value(), -2.); CHECK_DOUBLE_EQ(y.get_adjoint(), -1.); }
    { ForwardVar<double> y = ad::sin(seeded(0)); CHECK_DOUBLE_EQ(y.get_value(), 0.); CHECK_DOUBLE_EQ(y.get_adjoint(), 1.); }
    { ForwardVar<double> y = ad::cos(seeded(0)); CHECK_DOUBLE_EQ(y.get_value(), 1.); CHECK_DOUBLE_EQ(y.get_adjoint(), 0.); }
    { ForwardVar<double> y = ad::tan(seeded(0)); CHECK_DOUBLE_EQ(y.get_value(), 0.); CHECK_DOUBLE_EQ(y.get_adjoint(), 1.); }
    { ForwardVar<double> y = ad::asin(seeded(0)); CHECK_DOUBLE_EQ(y.get_value(), 0.); CHECK_DOUBLE_EQ(y.get_adjoint(), 1.); }
    { ForwardVar<double> y = ad::acos(seeded(0)); CHECK_DOUBLE_EQ(y.get_value(), M_PI / 2); CHECK_DOUBLE_EQ(y.get_adjoint(), -1.); }
    { ForwardVar<double> y = ad::atan(seeded(1)); CHECK_DOUBLE_EQ(y.get_value(), M_PI / 4); CHECK_DOUBLE_EQ(y.get_adjoint(), 0.5); }
    { ForwardVar<double> y = ad::exp(seeded(0)); CHECK_DOUBLE_EQ(y.get_value(), 1.); CHECK_DOUBLE_EQ(y.get_adjoint(), 1.); }
    { ForwardVar<double> y = ad::log(seeded(2)); CHECK_DOUBLE_EQ(y.get_value(), std::log(2)); CHECK_DOUBLE_EQ(y.get_adjoint(), 0.5); }
    { ForwardVar<double> y = ad::sqrt(seeded(4)); CHECK_DOUBLE_EQ(y.get_value(), 2.); CHECK_DOUBLE_EQ(y.get_adjoint(), 0.25); }
    { ForwardVar<double> y = ad::erf(seeded(4)); CHECK_DOUBLE_EQ(y.get_value(), 0.9999999845827420); CHECK_DOUBLE_EQ(y.get_adjoint(), 1.2698234671866558e-7); }
    // ---- binary
    { ForwardVar<double> x(4, 1), y(3, 1), r = x + y; CHECK_DOUBLE_EQ(r.get_value(), 7); CHECK_DOUBLE_EQ(r.get_adjoint(), 2); }
    { ForwardVar<double> x(4, 1), y(3, 1), r = x - y; CHECK_DOUBLE_EQ(r.get_value(), 1); CHECK_DOUBLE_EQ(r.get_adjoint(), 0); }
    { ForwardVar<double> x(4, 1), y(3, -1), r = x * y; CHECK_DOUBLE_EQ(r.get_value(), 12); CHECK_DOUBLE_EQ(r.get_adjoint(), -1); }
    { ForwardVar<double> x(4, 1), y(3, -1), r = x / y; CHECK_DOUBLE_EQ(r.get_value(), 4. / 3); CHECK_DOUBLE_EQ(r.get_adjoint(), 1. / 3 + 4. / 9); }
    // ---- operator+= (forward.hpp:160-164) and default construction (forward.hpp:82-84)
    {
        ForwardVar<double> acc;
        CHECK(acc.get_value() == 0 && acc.get_adjoint() == 0);
        acc += ForwardVar<double>(1.5, 0.25);
        acc += ForwardVar<double>(2.0, -1.0);
        CHECK_DOUBLE_EQ(acc.get_value(), 3.5);
        CHECK_DOUBLE_EQ(acc.get_adjoint(), -0.75);
    }
    // ---- README forward example shape: f(x, y) = sin(x) * exp(y) + x / y in direction (1, 0), then (0, 1)
    {
        const double xv = 0.7, yv = -1.3;
        ForwardVar<double> x(xv, 1), y(yv, 0);
        ForwardVar<double> f = ad::sin(x) * ad::exp(y) + x / y;
        CHECK_DOUBLE_EQ(f.get_value(), std::sin(xv) * std::exp(yv) + xv / yv);
        CHECK_DOUBLE_EQ(f.get_adjoint(), std::cos(xv) * std::exp(yv) + 1 / yv);
        x.set_adjoint(0); y.set_adjoint(1);
        f = ad::sin(x) * ad::exp(y) + x / y;
        CHECK_REL(f.get_adjoint(), std::sin(xv) * std::exp(yv) - xv / (yv * yv), 1e-15);
    }
    return report("test_forward");
}
