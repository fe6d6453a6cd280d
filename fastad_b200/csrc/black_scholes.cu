// black_scholes.cu — K1: batched Black-Scholes price + reverse-mode greeks.
//
// Replaces, for n independent options, the per-option loop of
//   auto e = ad::bind(black_scholes_option_price<cp>(S, K, sigma, tau, r, cache));
//   price = ad::autodiff(e);                       (example/black_scholes.cpp:16-39, 43-70)
// i.e. UnaryNode(Log, Erf, Exp, UnaryMinus) / BinaryNode(+,-,*,/) / EqNode / GlueNode
// feval + beval (unary.hpp:63-89, binary.hpp:100-136, eq.hpp:89-107, glue.hpp:67-86), with
// S, sigma and r as variables.  Each thread runs the forward pass and the adjoint sweep of
// BOTH expressions in registers (no tape, no caches in memory); the adjoint chain through the
// placeholders cache[0..2] is identical for call and put and is swept once.
//
// HBM layout: SoA.  5 input streams, 8 output streams, 104 algorithmic bytes per option.
#include <cstdlib>

#include "common.cuh"

namespace fadb {

struct BSOut { double* p[8]; };

struct BSRes {
    double call, put, cS, cSig, cR, pS, pSig, pR;
};

__device__ __forceinline__ BSRes bs_eval(double S, double K, double sigma, double tau, double r)
{
    constexpr double inv_sqrt2 = 0.70710678118654752440;
    // d/dx [0.5*(erf(x/sqrt2)+1)] = 0.5 * (2/sqrt(pi)) * exp(-(x/sqrt2)^2) / sqrt2
    // (Erf::bmap constant 1.1283791670955126, unary.hpp:264-271)
    constexpr double dphi_c = 0.5 * 1.1283791670955126 * inv_sqrt2;

    // ---- forward (feval) ----
    const double sqrt_tau = sqrt(tau);
    const double x = S / K;
    const double c0 = log(x);                                  // cache[0]
    const double sst = sigma * sqrt_tau;
    const double drift = (r + sigma * sigma / 2.) * tau;
    const double inv_sst = 1. / sst;
    const double c1 = (c0 + drift) * inv_sst;                  // cache[1]
    const double c2 = c1 - sst;                                // cache[2]
    const double u1 = c1 * inv_sqrt2, u2 = c2 * inv_sqrt2;
    const double e1 = erf(u1), e2 = erf(u2);
    const double Phi1 = 0.5 * (e1 + 1.), Phi2 = 0.5 * (e2 + 1.);
    const double Phi1n = 0.5 * (1. - e1), Phi2n = 0.5 * (1. - e2);   // Phi(-c1), Phi(-c2)
    const double disc = exp(-r * tau);
    const double PV = K * disc;

    BSRes o;
    o.call = Phi1 * S - Phi2 * PV;
    o.put = Phi2n * PV - Phi1n * S;

    // ---- backward (beval, seed 1).  d call / d c1 = S*phi(c1) = d put / d c1, and
    //      d call / d c2 = -PV*phi(c2) = d put / d c2: one shared sweep through cache[2..0].
    const double g_c2 = -(PV * (dphi_c * exp(-u2 * u2)));
    double g_c1 = S * (dphi_c * exp(-u1 * u1));
    // cache[2] = cache[1] - sst
    g_c1 += g_c2;
    double g_sst = -g_c2;
    // cache[1] = (cache[0] + drift) / sst
    const double g_num = g_c1 * inv_sst;
    g_sst += -g_c1 * c1 * inv_sst;
    // drift = (r + sigma^2/2) * tau ; sst = sigma * sqrt(tau)
    const double g_r_chain = g_num * tau;
    const double g_sig = g_r_chain * sigma + g_sst * sqrt_tau;
    // cache[0] = log(S / K)
    const double g_S_chain = g_num / x / K;
    // PV = K * exp(-r tau):  dPV/dr = -tau * PV
    const double dPV_dr = -tau * PV;

    o.cS = Phi1 + g_S_chain;
    o.cSig = g_sig;
    o.cR = g_r_chain - Phi2 * dPV_dr;
    o.pS = g_S_chain - Phi1n;
    o.pSig = g_sig;
    o.pR = g_r_chain + Phi2n * dPV_dr;
    return o;
}

template <bool ACC>
__device__ __forceinline__ void bs_store1(const BSOut& out, size_t i, const BSRes& o)
{
    out.p[0][i] = o.call;
    out.p[1][i] = o.put;
    const double g[6] = {o.cS, o.cSig, o.cR, o.pS, o.pSig, o.pR};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        if (ACC) out.p[2 + k][i] += g[k];
        else out.p[2 + k][i] = g[k];
    }
}

template <bool ACC, int MINB>
__global__ void __launch_bounds__(256, MINB)
bs_kernel(size_t n, const double* __restrict__ S, const double* __restrict__ K,
          const double* __restrict__ sigma, const double* __restrict__ tau,
          const double* __restrict__ r, BSOut out, bool vec_ok)
{
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    const size_t npair = vec_ok ? n / 2 : 0;
    for (size_t p = tid; p < npair; p += nthreads) {
        const double2 s2 = __ldg(reinterpret_cast<const double2*>(S) + p);
        const double2 k2 = __ldg(reinterpret_cast<const double2*>(K) + p);
        const double2 g2 = __ldg(reinterpret_cast<const double2*>(sigma) + p);
        const double2 t2 = __ldg(reinterpret_cast<const double2*>(tau) + p);
        const double2 r2 = __ldg(reinterpret_cast<const double2*>(r) + p);
        const BSRes a = bs_eval(s2.x, k2.x, g2.x, t2.x, r2.x);
        const BSRes b = bs_eval(s2.y, k2.y, g2.y, t2.y, r2.y);
        reinterpret_cast<double2*>(out.p[0])[p] = make_double2(a.call, b.call);
        reinterpret_cast<double2*>(out.p[1])[p] = make_double2(a.put, b.put);
        const double ga[6] = {a.cS, a.cSig, a.cR, a.pS, a.pSig, a.pR};
        const double gb[6] = {b.cS, b.cSig, b.cR, b.pS, b.pSig, b.pR};
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            double2* q = reinterpret_cast<double2*>(out.p[2 + k]) + p;
            double2 v = make_double2(ga[k], gb[k]);
            if (ACC) { const double2 old = *q; v.x += old.x; v.y += old.y; }
            *q = v;
        }
    }
    // tail (odd n) or unaligned buffers: scalar path
    for (size_t i = npair * 2 + tid; i < n; i += nthreads) {
        const BSRes o = bs_eval(S[i], K[i], sigma[i], tau[i], r[i]);
        bs_store1<ACC>(out, i, o);
    }
}

// one option per thread (64-bit coalesced accesses): half the live registers, twice the warps
template <bool ACC, int MINB>
__global__ void __launch_bounds__(256, MINB)
bs_kernel1(size_t n, const double* __restrict__ S, const double* __restrict__ K,
           const double* __restrict__ sigma, const double* __restrict__ tau,
           const double* __restrict__ r, BSOut out)
{
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    for (size_t i = tid; i < n; i += nthreads) {
        const BSRes o = bs_eval(__ldg(S + i), __ldg(K + i), __ldg(sigma + i), __ldg(tau + i), __ldg(r + i));
        bs_store1<ACC>(out, i, o);
    }
}

int launch_black_scholes(Context& c, cudaStream_t st, size_t n, const double* S, const double* K,
                         const double* sigma, const double* tau, const double* r,
                         double* const out[8], unsigned flags)
{
    BSOut o;
    bool vec_ok = (((uintptr_t)S | (uintptr_t)K | (uintptr_t)sigma | (uintptr_t)tau |
                    (uintptr_t)r) & 15u) == 0;
    for (int k = 0; k < 8; ++k) {
        o.p[k] = out[k];
        vec_ok = vec_ok && (((uintptr_t)out[k]) & 15u) == 0;
    }
    if (n == 0) return FADB_OK;
    const int threads = 256;
    static int variant = -1;
    if (variant < 0) { const char* e = getenv("FADB_BS_VARIANT"); variant = e ? atoi(e) : 5; }
#define FADB_BS2(MINB)                                                                                  \
    do {                                                                                                \
        if (flags & FADB_OVERWRITE_ADJ)                                                                 \
            bs_kernel<false, MINB><<<resident_grid(c, bs_kernel<false, MINB>, threads, 0, (n + 1) / 2, threads), \
                                     threads, 0, st>>>(n, S, K, sigma, tau, r, o, vec_ok);              \
        else                                                                                            \
            bs_kernel<true, MINB><<<resident_grid(c, bs_kernel<true, MINB>, threads, 0, (n + 1) / 2, threads),   \
                                    threads, 0, st>>>(n, S, K, sigma, tau, r, o, vec_ok);               \
    } while (0)
#define FADB_BS1(MINB)                                                                                  \
    do {                                                                                                \
        if (flags & FADB_OVERWRITE_ADJ)                                                                 \
            bs_kernel1<false, MINB><<<resident_grid(c, bs_kernel1<false, MINB>, threads, 0, n, threads), \
                                      threads, 0, st>>>(n, S, K, sigma, tau, r, o);                     \
        else                                                                                            \
            bs_kernel1<true, MINB><<<resident_grid(c, bs_kernel1<true, MINB>, threads, 0, n, threads),  \
                                     threads, 0, st>>>(n, S, K, sigma, tau, r, o);                      \
    } while (0)
    switch (variant) {
    case 1: FADB_BS2(3); break;
    case 2: FADB_BS2(4); break;
    case 3: FADB_BS1(4); break;
    case 4: FADB_BS1(6); break;
    case 0: FADB_BS2(2); break;
    default: FADB_BS1(3); break;   // measured best on B200: 1 option/thread, 60 regs, 4 CTAs/SM
    }
#undef FADB_BS2
#undef FADB_BS1
    c.launches++;
    FADB_CUDA(cudaGetLastError());
    return FADB_OK;
}

}  // namespace fadb

using namespace fadb;

extern "C" int fadb_black_scholes(size_t n, const double* S, const double* K,
                                  const double* sigma, const double* tau, const double* r,
                                  double* const out[8], unsigned flags)
{
    FADB_REQUIRE(out != nullptr, "fadb_black_scholes: null out table");
    if (n > 0) {
        FADB_REQUIRE(S && K && sigma && tau && r, "fadb_black_scholes: null input");
        for (int k = 0; k < 8; ++k) FADB_REQUIRE(out[k] != nullptr, "fadb_black_scholes: null output");
    }
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    return launch_black_scholes(*c, c->stream, n, S, K, sigma, tau, r, out, flags);
}

// Host-buffer entry: the call a user of the reference makes with plain host arrays.
// Chunks the batch, stages through pinned memory and overlaps H2D / kernel / D2H on
// three streams' worth of double buffering.
extern "C" int fadb_black_scholes_host(size_t n, const double* hS, const double* hK,
                                       const double* hsigma, const double* htau,
                                       const double* hr, double* const hout[8])
{
    FADB_REQUIRE(hout != nullptr, "fadb_black_scholes_host: null out table");
    if (n == 0) return FADB_OK;
    FADB_REQUIRE(hS && hK && hsigma && htau && hr, "fadb_black_scholes_host: null input");
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;

    constexpr int NBUF = 3;
    const size_t chunk = n < (size_t)(1u << 18) ? n : (size_t)(1u << 18);
    static thread_local double* dbuf[NBUF] = {nullptr, nullptr, nullptr};
    static thread_local size_t dcap = 0;
    static thread_local cudaStream_t st[NBUF];
    static thread_local cudaEvent_t done[NBUF];
    static thread_local cudaEvent_t fork_ev;
    static thread_local bool have_streams = false;
    if (!have_streams) {
        for (int b = 0; b < NBUF; ++b) {
            FADB_CUDA(cudaStreamCreateWithFlags(&st[b], cudaStreamNonBlocking));
            FADB_CUDA(cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming));
        }
        FADB_CUDA(cudaEventCreateWithFlags(&fork_ev, cudaEventDisableTiming));
        have_streams = true;
    }
    if (dcap < chunk) {
        for (int b = 0; b < NBUF; ++b) {
            if (dbuf[b]) FADB_CUDA(cudaFree(dbuf[b]));
            FADB_CUDA(cudaMalloc(&dbuf[b], sizeof(double) * 13 * chunk));
        }
        dcap = chunk;
    }
    // Register nothing: pageable host memory is accepted; cudaMemcpyAsync stages it.
    const double* hin[5] = {hS, hK, hsigma, htau, hr};
    // fork from / join back to the library stream so callers can bracket the call with events
    FADB_CUDA(cudaEventRecord(fork_ev, c->stream));
    for (int k = 0; k < NBUF; ++k) FADB_CUDA(cudaStreamWaitEvent(st[k], fork_ev, 0));
    int b = 0;
    for (size_t off = 0; off < n; off += chunk, b = (b + 1) % NBUF) {
        const size_t m = (n - off < chunk) ? n - off : chunk;
        double* base = dbuf[b];
        for (int k = 0; k < 5; ++k)
            FADB_CUDA(cudaMemcpyAsync(base + k * dcap, hin[k] + off, m * sizeof(double),
                                      cudaMemcpyHostToDevice, st[b]));
        double* dout[8];
        for (int k = 0; k < 8; ++k) dout[k] = base + (5 + k) * dcap;
        rc = launch_black_scholes(*c, st[b], m, base, base + dcap, base + 2 * dcap, base + 3 * dcap,
                                  base + 4 * dcap, dout, FADB_OVERWRITE_ADJ);
        if (rc) return rc;
        for (int k = 0; k < 8; ++k)
            FADB_CUDA(cudaMemcpyAsync(hout[k] + off, dout[k], m * sizeof(double),
                                      cudaMemcpyDeviceToHost, st[b]));
    }
    for (int k = 0; k < NBUF; ++k) {
        FADB_CUDA(cudaEventRecord(done[k], st[k]));
        FADB_CUDA(cudaStreamWaitEvent(c->stream, done[k], 0));
    }
    FADB_CUDA(cudaStreamSynchronize(c->stream));
    return FADB_OK;
}
