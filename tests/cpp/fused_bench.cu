// Small nvcc-compiled shared library used by bench.py to time the TYPE-FUSED path
// (include/fastad_b200/fused.cuh) on the same expressions the interpreter workloads use.
#include <fastad>
#include <fastad_b200/fused.cuh>

#include <memory>

using namespace ad;

template <class P> static auto Phi(const P& x) { return 0.5 * (ad::erf(x / std::sqrt(2.)) + 1.); }

namespace {
struct ChainState {
    VarView<double, vec> x;
    Var<double> w0{2.0}, w1{1.3};
    std::unique_ptr<core::FusedBind<decltype(ad::sum(-0.5 * ad::pow<2>((x - w0 * ad::constant_view((const double*)nullptr, 1)) / w1)))>> f;
};
ChainState* g_chain = nullptr;

auto bs_expr(VarView<double, vec>& S, VarView<double, vec>& sg, VarView<double, vec>& r, const double* K, const double* tau,
             const double* sq, size_t n, std::vector<VarView<double, vec>>& cache)
{
    auto Kc = ad::constant_view(K, n), tc = ad::constant_view(tau, n), sc = ad::constant_view(sq, n);
    auto common = (cache[0] = ad::log(S / Kc), cache[1] = (cache[0] + ((r + sg * sg / 2.) * tc)) / (sg * sc),
                   cache[2] = cache[1] - (sg * sc));
    auto PV = Kc * ad::exp(-r * tc);
    return (common, Phi(cache[1]) * S - Phi(cache[2]) * PV);
}
struct BSState {
    VarView<double, vec> S, sg, r;
    std::vector<VarView<double, vec>> cache;
    using expr_t = decltype(bs_expr(S, sg, r, nullptr, nullptr, nullptr, 1, cache));
    std::unique_ptr<core::FusedBind<expr_t>> f;
    const double* ones;
};
BSState* g_bs = nullptr;
}  // namespace

extern "C" {

// sum(-0.5 * pow<2>((x - w0*c) / w1)) over n device doubles
int fused_chain_setup(size_t n, double* x, double* x_adj, const double* c)
{
    try {
        delete g_chain;
        g_chain = new ChainState{VarView<double, vec>(x, x_adj, n)};
        auto cv = ad::constant_view(c, n);
        g_chain->f.reset(new std::decay_t<decltype(*g_chain->f)>(ad::sum(-0.5 * ad::pow<2>((g_chain->x - g_chain->w0 * cv) / g_chain->w1))));
        return 0;
    } catch (const std::exception& e) { std::fprintf(stderr, "%s\n", e.what()); return -1; }
}
int fused_chain_step(void)
{
    try { ad::autodiff_async(*g_chain->f, 1.0); return 0; } catch (...) { return -1; }
}

// batched Black-Scholes CALL expression with vec leaves and three vec placeholders; bufs = device arrays
// S, S_adj, sigma, sigma_adj, r, r_adj, K, tau, sqrt_tau, c0, c0_adj, c1, c1_adj, c2, c2_adj, ones
int fused_bs_setup(size_t n, double* const* b)
{
    try {
        delete g_bs;
        g_bs = new BSState{VarView<double, vec>(b[0], b[1], n), VarView<double, vec>(b[2], b[3], n), VarView<double, vec>(b[4], b[5], n), {}, nullptr, b[15]};
        for (int k = 0; k < 3; ++k) g_bs->cache.emplace_back(b[9 + 2 * k], b[10 + 2 * k], n);
        g_bs->f.reset(new core::FusedBind<BSState::expr_t>(bs_expr(g_bs->S, g_bs->sg, g_bs->r, b[6], b[7], b[8], n, g_bs->cache)));
        return 0;
    } catch (const std::exception& e) { std::fprintf(stderr, "%s\n", e.what()); return -1; }
}
int fused_bs_step(void)
{
    try { ad::autodiff(*g_bs->f, ad::device_seed{g_bs->ones}); return 0; } catch (...) { return -1; }
}
}
