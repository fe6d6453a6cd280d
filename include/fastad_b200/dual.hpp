// dual.hpp — forward-mode AD.
//
// ad::ForwardVar<T> is the dual number (value, tangent) of the reference's
// include/fastad_bits/forward/core/dualnum.hpp:10-50 and forward.hpp:78-167: same members
// (get_value / set_value / get_adjoint / set_adjoint, operator+=), same functions (unary minus,
// sin, cos, tan, asin, acos, atan, exp, log, sqrt, erf) and operators (+ - * /), and the same
// arithmetic in the same order, so host results are bit-identical to the reference's for equal
// libm results.
//
// B200-native part: every function is __host__ __device__ under nvcc, so a ForwardVar is two
// registers inside a kernel; fastad_b200/forward_batch.cuh sweeps a user functor over a batch of
// independent instances (one per thread, SoA values and tangents in HBM) — the forward-mode
// counterpart of the batched reverse-mode kernels.  Under g++ it is what it is in the reference:
// an eager value type, no graph, no device.
//
// Reference behaviour kept on purpose: ad::erf's tangent is (2/sqrt(pi)) exp(-x^2) WITHOUT the
// factor x.get_adjoint() (forward.hpp:131-135; forward_unittest.cpp:108-115 seeds it with 1).
#pragma once
#include <math.h>
#include <cmath>

#if defined(__CUDACC__)
#define FADB_HD __host__ __device__ __forceinline__
#else
#define FADB_HD inline
#endif

namespace ad {
namespace core {

template <class T>
struct DualNum {                                                     // dualnum.hpp:10-50
    using value_type = T;
    FADB_HD DualNum(T value, T tangent) : v_(value), t_(tangent) {}
    FADB_HD value_type& get_value() { return v_; }
    FADB_HD const value_type& get_value() const { return v_; }
    FADB_HD value_type& set_value(value_type x) { return v_ = x; }
    FADB_HD value_type& get_adjoint() { return t_; }
    FADB_HD const value_type& get_adjoint() const { return t_; }
    FADB_HD value_type& set_adjoint(value_type x) { return t_ = x; }

private:
    T v_, t_;
};

template <class T>
struct ADForward : DualNum<T> {                                      // forward.hpp:78-93
    using data_t = DualNum<T>;
    FADB_HD ADForward() : data_t(0, 0) {}
    FADB_HD ADForward(T value, T tangent = 0) : data_t(value, tangent) {}
    FADB_HD ADForward& operator+=(const ADForward& o)                // forward.hpp:160-164
    {
        const T v = this->get_value() + o.get_value(), t = this->get_adjoint() + o.get_adjoint();
        this->set_value(v);
        this->set_adjoint(t);
        return *this;
    }
};

// forward.hpp:143-157 — found by ADL on ADForward, like the reference's
template <class T>
FADB_HD ADForward<T> operator+(const ADForward<T>& a, const ADForward<T>& b)
{
    return {a.get_value() + b.get_value(), a.get_adjoint() + b.get_adjoint()};
}
template <class T>
FADB_HD ADForward<T> operator-(const ADForward<T>& a, const ADForward<T>& b)
{
    return {a.get_value() - b.get_value(), a.get_adjoint() - b.get_adjoint()};
}
template <class T>
FADB_HD ADForward<T> operator*(const ADForward<T>& a, const ADForward<T>& b)
{
    const T av = a.get_value(), bv = b.get_value();
    return {av * bv, av * b.get_adjoint() + a.get_adjoint() * bv};
}
template <class T>
FADB_HD ADForward<T> operator/(const ADForward<T>& a, const ADForward<T>& b)
{
    const T av = a.get_value(), bv = b.get_value();
    return {av / bv, (a.get_adjoint() * bv - av * b.get_adjoint()) / (bv * bv)};
}

// forward.hpp:104 (declared in ad:: there; here in core:: so that ADL finds it from any namespace)
template <class T>
FADB_HD ADForward<T> operator-(const ADForward<T>& a)
{
    return {-a.get_value(), -a.get_adjoint()};
}

}  // namespace core

template <class T>
using ForwardVar = core::ADForward<T>;                               // forward.hpp:97-98

// Unary functions, forward.hpp:104-135.  `dx` is the incoming tangent.
#define FADB_DUAL_UNARY(NAME, SETUP, VALUE, TANGENT)                 \
    template <class T>                                               \
    FADB_HD core::ADForward<T> NAME(const core::ADForward<T>& u)     \
    {                                                                \
        const T x = u.get_value(), dx = u.get_adjoint();             \
        (void)dx;                                                    \
        SETUP;                                                       \
        return {VALUE, TANGENT};                                     \
    }
FADB_DUAL_UNARY(sin, , ::sin(x), ::cos(x) * dx)
FADB_DUAL_UNARY(cos, , ::cos(x), -::sin(x) * dx)
FADB_DUAL_UNARY(tan, const T sec = 1 / ::cos(x), ::tan(x), sec * sec * dx)
FADB_DUAL_UNARY(asin, , ::asin(x), dx / (::sqrt(1 - x * x)))
FADB_DUAL_UNARY(acos, , ::acos(x), -dx / (::sqrt(1 - x * x)))
FADB_DUAL_UNARY(atan, , ::atan(x), dx / (1 + x * x))
FADB_DUAL_UNARY(exp, const T e = ::exp(x), e, e * dx)
FADB_DUAL_UNARY(log, , ::log(x), dx / x)
FADB_DUAL_UNARY(sqrt, const T s = ::sqrt(x), s, dx / (2 * s))
FADB_DUAL_UNARY(erf, const T xx = x * x, ::erf(x), T(1.1283791670955126) * ::exp(-xx))   // sic: no dx
#undef FADB_DUAL_UNARY

}  // namespace ad
