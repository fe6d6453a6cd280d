// tiny assertion harness for the C++ host-layer tests (gtest is not in this image)
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>

static int g_checks = 0, g_failures = 0;

inline double ulp_diff(double a, double b)
{
    if (a == b) return 0;
    if (std::isnan(a) || std::isnan(b) || std::isinf(a) || std::isinf(b)) return std::numeric_limits<double>::infinity();
    double sp = std::nextafter(std::fabs(b), std::numeric_limits<double>::infinity()) - std::fabs(b);
    return std::fabs(a - b) / sp;
}
#define CHECK(cond)                                                                  \
    do {                                                                             \
        ++g_checks;                                                                  \
        if (!(cond)) { ++g_failures; std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); } \
    } while (0)
// EXPECT_DOUBLE_EQ of the reference = 4 ulp
#define CHECK_DOUBLE_EQ(a, b)                                                        \
    do {                                                                             \
        ++g_checks;                                                                  \
        double _a = (a), _b = (b);                                                   \
        if (!(ulp_diff(_a, _b) <= 4)) { ++g_failures; std::printf("FAIL %s:%d  %s = %.17g  vs  %s = %.17g\n", __FILE__, __LINE__, #a, _a, #b, _b); } \
    } while (0)
#define CHECK_NEAR(a, b, tol)                                                        \
    do {                                                                             \
        ++g_checks;                                                                  \
        double _a = (a), _b = (b);                                                   \
        if (!(std::fabs(_a - _b) <= (tol))) { ++g_failures; std::printf("FAIL %s:%d  %s = %.17g  vs  %s = %.17g\n", __FILE__, __LINE__, #a, _a, #b, _b); } \
    } while (0)
// arguments are evaluated exactly once (they may be ad::autodiff calls, which accumulate adjoints)
#define CHECK_REL(a, b, rel)                                                         \
    do {                                                                             \
        ++g_checks;                                                                  \
        double _a = (a), _b = (b);                                                   \
        if (!(std::fabs(_a - _b) <= (rel) * std::fabs(_b))) { ++g_failures; std::printf("FAIL %s:%d  %s = %.17g  vs  %s = %.17g\n", __FILE__, __LINE__, #a, _a, #b, _b); } \
    } while (0)
inline int report(const char* name)
{
    std::printf("%s: %d checks, %d failures\n", name, g_checks, g_failures);
    return g_failures ? 1 : 0;
}
