// device_math.cuh — device forward/backward maps of the elementwise nodes (shared by the
// library kernels in fastad_b200/csrc and by the type-fused kernels of fused.cuh).
// fmap/bmap of UnaryNode (reverse/core/unary.hpp:190-296), fmap/blmap/brmap of BinaryNode
// (binary.hpp:265-338) and PowNode<n> (pow.hpp:20-56, 91-162), each written in the operation
// order of the reference so that per-instance results agree to the last few ulp (CUDA's
// libdevice exp/log/erf differ from glibc's by <= 1-2 ulp).
#pragma once
#ifndef __CUDACC_RTC__   // NVRTC (specialised stage kernels, csrc/jit.cu) gets these files pre-concatenated
#include <cuda_runtime.h>

#include <cmath>

#include "../fastad_b200.h"
#include "fastmath.cuh"
#else
#ifndef INFINITY
#define INFINITY __longlong_as_double(0x7ff0000000000000LL)
#endif
#ifndef NAN
#define NAN __longlong_as_double(0x7ff8000000000000LL)
#endif
#endif

namespace fadb {

template <int OP> struct Unary;

#define FADB_UNARY(OP, FEXPR, BEXPR)                                                      \
    template <> struct Unary<OP> {                                                        \
        __device__ __forceinline__ static double f(double x) { return FEXPR; }            \
        __device__ __forceinline__ static double b(double seed, double x, double f)       \
        { (void)x; (void)f; return BEXPR; }                                               \
    };

FADB_UNARY(FADB_NEG, -x, -seed)
FADB_UNARY(FADB_SIN, sin(x), seed * cos(x))
FADB_UNARY(FADB_COS, cos(x), -seed * sin(x))
FADB_UNARY(FADB_TAN, tan(x), seed / (cos(x) * cos(x)))
FADB_UNARY(FADB_ASIN, asin(x), seed / sqrt(1. - x * x))
FADB_UNARY(FADB_ACOS, acos(x), -(seed / sqrt(1. - x * x)))
FADB_UNARY(FADB_ATAN, atan(x), seed / (1. + x * x))
// exp / log: constant-bank polynomials of fastmath.cuh (<= 1 ulp, validated against glibc on the
// host) on the ordinary range, libdevice for everything else (NaN, inf, overflow, subnormal
// results, non-positive log arguments) so that IEEE special-case behaviour is unchanged.
__device__ __forceinline__ double exp_guarded(double x)
{
    return (fabs(x) <= 700.0) ? fm::exp_(x) : exp(x);
}
__device__ __forceinline__ double log_guarded(double x)
{
    const unsigned long long ux = (unsigned long long)__double_as_longlong(x);
    // normal, positive, finite
    return (ux - 0x0010000000000000ull < 0x7fe0000000000000ull) ? fm::log_(x) : log(x);
}
FADB_UNARY(FADB_EXP, exp_guarded(x), seed * f)
// Division by Newton reciprocal + residual correction (fm::div_, <= 1 ulp) when both operands are
// "ordinary", IEEE division otherwise.  Ordinary = biased exponent in [512, 1534], i.e. 2^-511 <= |v| <
// 2^512 (or a == 0): the quotient then lies in [2^-1023, 2^1023) and neither it nor the residual can
// overflow or go subnormal.  The window test is two integer instructions on the high word (the
// earlier form, four FP64 compares against 1e+-290, cost 5 DSETP + 8 UMOV per division and let
// |a/d| > DBL_MAX through to the Newton path, which returns NaN instead of inf there).
__device__ __forceinline__ bool ordinary_(double v)
{
    return ((((unsigned)__double2hiint(v)) >> 20) & 0x7ffu) - 512u < 1023u;
}
__device__ __forceinline__ double div_guarded(double a, double d)
{
    if (ordinary_(d) && (ordinary_(a) || a == 0.0)) return fm::div_(a, d);
    return a / d;
}
// a1 / d and a2 / d with one reciprocal (Div::blmap / brmap share the denominator, binary.hpp:320-338)
__device__ __forceinline__ void div2_guarded(double a1, double a2, double d, double& q1, double& q2)
{
    if (ordinary_(d) && (ordinary_(a1) || a1 == 0.0) && (ordinary_(a2) || a2 == 0.0)) {
        const double y = fm::rcp_(d);
        q1 = fm::div_(a1, d, y);
        q2 = fm::div_(a2, d, y);
    } else {
        q1 = a1 / d;
        q2 = a2 / d;
    }
}
FADB_UNARY(FADB_LOG, log_guarded(x), div_guarded(seed, x))
FADB_UNARY(FADB_SQRT, sqrt(x), 0.5 * seed / f)
FADB_UNARY(FADB_ERF, erf(x), 1.1283791670955126 * seed * exp_guarded(-x * x))
FADB_UNARY(FADB_SINH, sinh(x), seed * cosh(x))
FADB_UNARY(FADB_COSH, cosh(x), seed * sinh(x))
FADB_UNARY(FADB_TANH, tanh(x), seed * (1 - f * f))
FADB_UNARY(FADB_MOCK2X, 2 * x, seed * 2)
FADB_UNARY(FADB_SQUARE, x * x, seed * 2. * x)      // NormNode, norm.hpp:47-57
#undef FADB_UNARY

// Sigmoid: the reference evaluates exp(-x) three times in bmap (unary.hpp:273-278); the value
// is the same each time, so it is computed once.
template <> struct Unary<FADB_SIGMOID> {
    __device__ __forceinline__ static double f(double x) { return div_guarded(1., 1 + exp_guarded(-x)); }
    __device__ __forceinline__ static double b(double seed, double x, double)
    {
        const double e = exp_guarded(-x);
        return div_guarded(seed * e, (e + 1) * (e + 1));
    }
};

// PowFunc (pow.hpp:20-56): repeated squaring for n >= 0, reciprocal for n < 0.
__device__ __forceinline__ double pow_int(long n, double base)
{
    if (n == 2) return base * base;
    if (n == 3) return base * (base * base);
    bool neg = n < 0;
    if (neg) {
        if (base == 0) return INFINITY;
        base = 1. / base;
        n = -n;
    }
    // evaluate PowFunc<n> bottom-up: same multiplication tree as the recursive template
    // (n even: t*t with t = PowFunc<n/2>; n odd: base * PowFunc<n-1>)
    double result = 1.;
    int bits[64];
    int nb = 0;
    for (long m = n; m > 0;) {
        if (m & 1) { bits[nb++] = 1; m -= 1; }
        else { bits[nb++] = 0; m >>= 1; }
    }
    for (int k = nb - 1; k >= 0; --k) result = bits[k] ? base * result : result * result;
    return result;
}

// PowNode<n>::beval seed for the child (pow.hpp:101-162)
__device__ __forceinline__ double pow_bmap(long n, double seed, double x, double f)
{
    if (n == 0) return 0.;
    if (n == 1) return seed;
    if (n > 1) return (double)n * seed * (x == 0 ? 0 : f / x);
    return x == 0 ? -INFINITY : (double)n * seed * f / x;
}

// run-time dispatch used by the tape interpreter (expr.cu)
__device__ __forceinline__ double unary_f_rt(int op, double x)
{
    switch (op) {
#define C(OP) case OP: return Unary<OP>::f(x);
        C(FADB_NEG) C(FADB_SIN) C(FADB_COS) C(FADB_TAN) C(FADB_ASIN) C(FADB_ACOS) C(FADB_ATAN)
        C(FADB_EXP) C(FADB_LOG) C(FADB_SQRT) C(FADB_ERF) C(FADB_SIGMOID) C(FADB_SINH) C(FADB_COSH)
        C(FADB_TANH) C(FADB_MOCK2X) C(FADB_SQUARE)
#undef C
    }
    return NAN;
}
__device__ __forceinline__ double unary_b_rt(int op, double seed, double x, double f)
{
    switch (op) {
#define C(OP) case OP: return Unary<OP>::b(seed, x, f);
        C(FADB_NEG) C(FADB_SIN) C(FADB_COS) C(FADB_TAN) C(FADB_ASIN) C(FADB_ACOS) C(FADB_ATAN)
        C(FADB_EXP) C(FADB_LOG) C(FADB_SQRT) C(FADB_ERF) C(FADB_SIGMOID) C(FADB_SINH) C(FADB_COSH)
        C(FADB_TANH) C(FADB_MOCK2X) C(FADB_SQUARE)
#undef C
    }
    return NAN;
}

__device__ __forceinline__ double binary_f_rt(int op, double x, double y)
{
    switch (op) {
    case FADB_ADD: return x + y;
    case FADB_SUB: return x - y;
    case FADB_MUL: return x * y;
    case FADB_DIV: return x / y;
    case FADB_MOCKB: return x - 2 * y;
    }
    return NAN;
}
__device__ __forceinline__ double binary_bl_rt(int op, double s, double x, double y, double f)
{
    (void)x; (void)f;
    switch (op) {
    case FADB_ADD: case FADB_SUB: case FADB_MOCKB: return s;
    case FADB_MUL: return s * y;
    case FADB_DIV: return s / y;
    }
    return NAN;
}
__device__ __forceinline__ double binary_br_rt(int op, double s, double x, double y, double f)
{
    switch (op) {
    case FADB_ADD: return s;
    case FADB_SUB: return -s;
    case FADB_MUL: return s * x;
    case FADB_DIV: return -s * f / y;
    case FADB_MOCKB: return -2. * s;
    }
    return NAN;
}

}  // namespace fadb
