// GPU, nvcc: batched forward-mode sweeps (include/fastad_b200/forward_batch.cuh) against
//  (1) the same functor evaluated with the host ForwardVar (same formulas, glibc vs libdevice: 4 ulp
//      plus the expression's conditioning floor), and
//  (2) the reverse-mode Black-Scholes kernel: a forward sweep in direction e_S must give the delta.
#include <fastad>
#include <fastad_b200/forward_batch.cuh>
#include <random>
#include <vector>

#include "check.hpp"

using FV = ad::ForwardVar<double>;

struct Smooth {   // f(x, y) = sin(x) exp(y) / sqrt(x) + log(x) atan(y) - tan(y) cos(x)
    __host__ __device__ FV operator()(const FV& x, const FV& y) const
    {
        return ad::sin(x) * ad::exp(y) / ad::sqrt(x) + ad::log(x) * ad::atan(y) - ad::tan(y) * ad::cos(x);
    }
};
struct CallPrice {   // example/black_scholes.cpp:16-39 written on dual numbers; sigma, tau, r, K carry tangent 0
    __host__ __device__ FV operator()(const FV& S, const FV& K, const FV& sg, const FV& tau, const FV& r) const
    {
        const FV half(0.5), one(1.0), rs2(0.70710678118654752440);
        const FV sq = ad::sqrt(tau);
        const FV d1 = (ad::log(S / K) + (r + sg * sg * half) * tau) / (sg * sq);
        const FV d2 = d1 - sg * sq;
        // Phi via erf: the reference's erf tangent drops dx (forward.hpp:131-135), so carry it by hand
        auto Phi = [&](const FV& d) {
            const FV e = ad::erf(d * rs2);
            return FV(0.5 * (e.get_value() + 1.0), 0.5 * e.get_adjoint() * (d * rs2).get_adjoint());
        };
        return Phi(d1) * S - Phi(d2) * K * ad::exp(-r * tau);
    }
};

template <class T> static T* dev_copy(const std::vector<T>& h)
{
    void* d = nullptr;
    if (fadb_malloc(h.size() * sizeof(T), &d) || fadb_h2d(d, h.data(), h.size() * sizeof(T))) { std::printf("%s\n", fadb_last_error()); std::exit(2); }
    return static_cast<T*>(d);
}

int main()
{
    if (fadb_init(0) != 0) { std::printf("no CUDA device: %s\n", fadb_last_error()); return 2; }
    std::mt19937_64 gen(20261020);
    // ---- (1) two-input smooth function, per-instance tangents for x, constant direction for y
    {
        const size_t n = 70001;
        std::vector<double> x(n), y(n), tx(n), ov(n), ot(n);
        std::uniform_real_distribution<double> U(0.3, 1.2), D(-1, 1);
        for (size_t i = 0; i < n; ++i) { x[i] = U(gen); y[i] = D(gen); tx[i] = D(gen); }
        double *dx = dev_copy(x), *dy = dev_copy(y), *dtx = dev_copy(tx), *dov = dev_copy(ov), *dot = dev_copy(ot);
        const double* in_val[2] = {dx, dy};
        const double* in_tan[2] = {dtx, nullptr};
        const double dir[2] = {0.0, 0.5};
        CHECK(ad::forward_batch<2>(n, Smooth{}, in_val, in_tan, dir, dov, dot) == 0);
        CHECK(fadb_d2h(ov.data(), dov, n * 8) == 0 && fadb_d2h(ot.data(), dot, n * 8) == 0);
        double worst_v = 0, worst_t = 0;
        for (size_t i = 0; i < n; ++i) {
            const FV h = Smooth{}(FV(x[i], tx[i]), FV(y[i], 0.5));
            // a dozen terms of magnitude <= ~4, each within 2 ulp between glibc and libdevice
            worst_v = std::max(worst_v, std::fabs(ov[i] - h.get_value()) / std::max(std::fabs(h.get_value()), 4.0));
            worst_t = std::max(worst_t, std::fabs(ot[i] - h.get_adjoint()) / std::max(std::fabs(h.get_adjoint()), 4.0));
        }
        CHECK(worst_v <= 4e-15);
        CHECK(worst_t <= 1e-14);
        // n = 0 is a no-op, null out_val is allowed
        CHECK(ad::forward_batch<2>(0, Smooth{}, in_val, in_tan, dir, dov, dot) == 0);
        CHECK(ad::forward_batch<2>(n, Smooth{}, in_val, in_tan, dir, nullptr, dot) == 0);
        std::vector<double> ot2(n);
        CHECK(fadb_d2h(ot2.data(), dot, n * 8) == 0);
        CHECK(std::memcmp(ot.data(), ot2.data(), n * 8) == 0);   // deterministic
        for (double* p : {dx, dy, dtx, dov, dot}) fadb_free(p);
    }
    // ---- (2) forward sweep in direction e_S == reverse-mode delta of the Black-Scholes kernel
    {
        const size_t n = 1 << 16;
        std::vector<double> S(n), K(n), sg(n), tau(n), r(n);
        std::uniform_real_distribution<double> US(50, 150), Usg(0.05, 0.6), Ut(0.05, 2), Ur(0, 0.08);
        for (size_t i = 0; i < n; ++i) { S[i] = US(gen); K[i] = US(gen); sg[i] = Usg(gen); tau[i] = Ut(gen); r[i] = Ur(gen); }
        double *dS = dev_copy(S), *dK = dev_copy(K), *dsg = dev_copy(sg), *dtau = dev_copy(tau), *dr = dev_copy(r);
        std::vector<double> z(n, 0.0);
        double* out[8];
        for (auto& p : out) p = dev_copy(z);
        CHECK(fadb_black_scholes(n, dS, dK, dsg, dtau, dr, out, FADB_OVERWRITE_ADJ) == 0);
        double *fv = dev_copy(z), *ft = dev_copy(z);
        const double* in_val[5] = {dS, dK, dsg, dtau, dr};
        const double* in_tan[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
        std::vector<double> call(n), rev(n), fwdv(n), fwdt(n);
        CHECK(fadb_d2h(call.data(), out[0], n * 8) == 0);
        for (int k = 0; k < 3; ++k) {   // directions e_S, e_sigma, e_r  <->  out[2], out[3], out[4]
            const double dir[5] = {k == 0 ? 1.0 : 0.0, 0.0, k == 1 ? 1.0 : 0.0, 0.0, k == 2 ? 1.0 : 0.0};
            CHECK(ad::forward_batch<5>(n, CallPrice{}, in_val, in_tan, dir, fv, ft) == 0);
            CHECK(fadb_d2h(fwdv.data(), fv, n * 8) == 0 && fadb_d2h(fwdt.data(), ft, n * 8) == 0);
            CHECK(fadb_d2h(rev.data(), out[2 + k], n * 8) == 0);
            double wv = 0, wt = 0;
            for (size_t i = 0; i < n; ++i) {
                // prices and greeks subtract two terms of size ~S: conditioning floor 1e-13 * S
                wv = std::max(wv, std::fabs(fwdv[i] - call[i]) / (std::fabs(call[i]) + 1e-3 * S[i]));
                wt = std::max(wt, std::fabs(fwdt[i] - rev[i]) / (std::fabs(rev[i]) + 1e-3 * S[i]));
            }
            CHECK(wv <= 1e-11);
            CHECK(wt <= 1e-10);
        }
        for (double* p : {dS, dK, dsg, dtau, dr, fv, ft}) fadb_free(p);
        for (auto p : out) fadb_free(p);
    }
    return report("test_forward_batch");
}
