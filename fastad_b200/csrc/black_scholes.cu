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
#include <algorithm>
#include <chrono>
#include <cstdlib>

#include "common.cuh"
#include "tma_util.cuh"
#include "../../include/fastad_b200/fastmath.cuh"

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

// Same forward pass and adjoint sweep with the constant-bank / Newton math of fastmath.cuh:
// one shared reciprocal per division, exp(-u^2) shared between Phi(d) and its derivative.
// PW: Phi through the 32-interval degree-7 table (fm::phi_pair_pw, `tab` = its shared-memory copy)
// instead of the single degree-27 polynomial: 40 fewer FP64 instructions per option
template <bool PW = false>
__device__ __forceinline__ BSRes bs_eval_fast(double S, double K, double sigma, double tau, double r, const double* tab = nullptr)
{
    constexpr double inv_sqrt2 = 0.70710678118654752440;
    constexpr double dphi_c = 0.5 * 1.1283791670955126 * inv_sqrt2;

    // ---- forward (feval) ----
    const double sqrt_tau = fm::sqrt_(tau);
    const double inv_K = fm::rcp_(K);
    const double x = fm::div_(S, K, inv_K);                    // S / K
    const double c0 = PW ? fm::log_t(x, tab + 288) : fm::log_(x);   // cache[0]
    const double sst = sigma * sqrt_tau;
    const double drift = (r + sigma * sigma * 0.5) * tau;
    const double inv_sst = fm::rcp_(sst);
    const double c1 = fm::div_(c0 + drift, sst, inv_sst);      // cache[1]
    const double c2 = c1 - sst;                                // cache[2]
    double Phi1, Phi1n, E1, Phi2, Phi2n, E2;
    if (PW) {
        fm::phi_pair_pw(c1, tab, Phi1, Phi1n, E1);
        fm::phi_pair_pw(c2, tab, Phi2, Phi2n, E2);
    } else {
        fm::phi_pair(c1, Phi1, Phi1n, E1);
        fm::phi_pair(c2, Phi2, Phi2n, E2);
    }
    const double PV = K * (PW ? fm::exp_t(-r * tau, tab + 256) : fm::exp_(-r * tau));

    BSRes o;
    o.call = Phi1 * S - Phi2 * PV;
    o.put = Phi2n * PV - Phi1n * S;

    // ---- backward (beval, seed 1), shared sweep through cache[2..0] ----
    const double g_c2 = -(PV * (dphi_c * E2));
    double g_c1 = S * (dphi_c * E1);
    g_c1 += g_c2;
    double g_sst = -g_c2;
    const double g_num = g_c1 * inv_sst;
    g_sst += -g_c1 * c1 * inv_sst;
    const double g_r_chain = g_num * tau;
    const double g_sig = g_r_chain * sigma + g_sst * sqrt_tau;
    const double g_S_chain = g_num * fm::rcp_(x) * inv_K;      // g_num / x / K
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

// one option per thread (64-bit coalesced accesses): half the live registers, twice the warps
template <bool ACC>
__global__ void __launch_bounds__(256, 3)
bs_kernel_libdevice(size_t n, const double* __restrict__ S, const double* __restrict__ K,
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

template <bool ACC>
__global__ void __launch_bounds__(256, 3)
bs_kernel_fast(size_t n, const double* __restrict__ S, const double* __restrict__ K,
               const double* __restrict__ sigma, const double* __restrict__ tau,
               const double* __restrict__ r, BSOut out)
{
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    for (size_t i = tid; i < n; i += nthreads) {
        const BSRes o = bs_eval_fast(__ldg(S + i), __ldg(K + i), __ldg(sigma + i), __ldg(tau + i), __ldg(r + i));
        bs_store1<ACC>(out, i, o);
    }
}

// software-pipelined variant: the next option's inputs are loaded before the current one is
// evaluated, so the ~800-cycle load latency overlaps the ~2000-cycle evaluation of the same thread
// EARLY (batches 1.. of fadb_black_scholes_batches, never the first kernel of a call): the previous grid on the stream is
// this same kernel working on ANOTHER operand set, so nothing this grid reads or writes depends on it.  The grid then
// skips the dependency wait at its start -- its CTAs begin streaming as soon as the previous grid's CTAs free their
// slots, instead of idling until the slowest CTA of the previous grid has retired (the ~1.5 us bubble per launch that
// separates the kernel at 2^20 options from its own steady state at 2^24) -- and executes the wait at its END, so that
// its completion still implies the completion (and visibility) of every grid before it: what follows on the stream
// sees exactly the ordinary stream order.
// The batch AHEAD of such a chain (PDL_HEAD) waits first and only then lets its dependents launch: the early grids behind
// it can therefore never run before whatever preceded the call on the stream (possibly the producer of the inputs, possibly
// itself a kernel that triggers early) has completed and is visible.
enum : int { PDL_PLAIN = 0, PDL_EARLY = 1, PDL_HEAD = 2 };
template <bool ACC, bool PW = false, int MINB = 2, int THREADS = 256, int PDL = PDL_PLAIN>
__global__ void __launch_bounds__(THREADS, MINB)
bs_kernel_fast_pf(size_t n, const double* __restrict__ S, const double* __restrict__ K,
                  const double* __restrict__ sigma, const double* __restrict__ tau,
                  const double* __restrict__ r, BSOut out)
{
    __shared__ double s_tab[PW ? 675 : 1];      // erfcx pieces [0, 256) + 2^(j/32) [256, 288) + log table [288, 675)
    if (PW) {            // constant memory: written before any launch, safe to read ahead of the dependency sync
        for (int i = threadIdx.x; i < 256; i += THREADS) s_tab[i] = fm::kErfcx32_d[i];
        if (threadIdx.x < 32) s_tab[256 + threadIdx.x] = fm::kExp2_32_d[threadIdx.x];
        for (int i = threadIdx.x; i < 387; i += THREADS) s_tab[288 + i] = fm::kLogT_d[i];
        __syncthreads();
    }
    if (PDL == PDL_EARLY) cudaTriggerProgrammaticLaunchCompletion();
    else if (PDL == PDL_HEAD) { cudaGridDependencySynchronize(); cudaTriggerProgrammaticLaunchCompletion(); }
    else pdl_prologue();      // common.cuh: no memory access before the previous grid on the stream has completed
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    size_t i = tid;
    if (i < n) {
        double s = __ldg(S + i), k = __ldg(K + i), g = __ldg(sigma + i), t = __ldg(tau + i), rr = __ldg(r + i);
        while (true) {
            const size_t j = i + nthreads;
            const bool more = j < n;
            double s2 = 0, k2 = 1, g2 = 1, t2 = 1, r2 = 0;
            if (more) { s2 = __ldg(S + j); k2 = __ldg(K + j); g2 = __ldg(sigma + j); t2 = __ldg(tau + j); r2 = __ldg(r + j); }
            const BSRes o = bs_eval_fast<PW>(s, k, g, t, rr, s_tab);
            bs_store1<ACC>(out, i, o);
            if (!more) break;
            i = j; s = s2; k = k2; g = g2; t = t2; rr = r2;
        }
    }
    if (PDL == PDL_EARLY && threadIdx.x == 0) cudaGridDependencySynchronize();     // the CTA retires only after the previous grid has
}

// Host-link variant (fadb_black_scholes_host on page-locked arrays, FADB_BS_HOST_WIDE=W): W = 2 or 4 consecutive options per
// thread, every array moved with ONE 128- / 256-bit access per thread -- a warp instruction then covers 512 B / 1 KB of
// contiguous host memory, so that the write path can form full-size PCIe payloads (with one option per thread a warp store is
// two separate 128-byte lines).  Overwrite mode only (the host entry's); same bs_eval_fast, same bits.
template <int W>
__global__ void __launch_bounds__(256, 1)
bs_kernel_host_wide(size_t n, const double* __restrict__ S, const double* __restrict__ K, const double* __restrict__ sigma,
                    const double* __restrict__ tau, const double* __restrict__ r, BSOut out)
{
    __shared__ double s_tab[288 + 387];
    for (int i = threadIdx.x; i < 256; i += 256) s_tab[i] = fm::kErfcx32_d[i];
    if (threadIdx.x < 32) s_tab[256 + threadIdx.x] = fm::kExp2_32_d[threadIdx.x];
    for (int i = threadIdx.x; i < 387; i += 256) s_tab[288 + i] = fm::kLogT_d[i];
    __syncthreads();
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    const size_t nw = n / W;
    for (size_t q = tid; q < nw; q += nthreads) {
        const size_t i = q * W;
        double in[5][W], res[8][W];
        const double* src[5] = {S, K, sigma, tau, r};
#pragma unroll
        for (int a = 0; a < 5; ++a) {
            if (W == 4) { const double4_t v = ldg256_stream(src[a] + i); for (int e = 0; e < 4; ++e) in[a][e] = v.v[e]; }
            else { const double2 v = __ldg(reinterpret_cast<const double2*>(src[a] + i)); in[a][0] = v.x; in[a][W - 1] = v.y; }
        }
#pragma unroll
        for (int e = 0; e < W; ++e) {
            const BSRes o = bs_eval_fast<true>(in[0][e], in[1][e], in[2][e], in[3][e], in[4][e], s_tab);
            res[0][e] = o.call; res[1][e] = o.put; res[2][e] = o.cS; res[3][e] = o.cSig; res[4][e] = o.cR;
            res[5][e] = o.pS; res[6][e] = o.pSig; res[7][e] = o.pR;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (W == 4) { double4_t v; for (int e = 0; e < 4; ++e) v.v[e] = res[k][e]; st256(out.p[k] + i, v); }
            else *reinterpret_cast<double2*>(out.p[k] + i) = make_double2(res[k][0], res[k][W - 1]);
        }
    }
    for (size_t i = nw * W + tid; i < n; i += nthreads) {
        const BSRes o = bs_eval_fast<true>(__ldg(S + i), __ldg(K + i), __ldg(sigma + i), __ldg(tau + i), __ldg(r + i), s_tab);
        bs_store1<false>(out, i, o);
    }
}

// Diagnostic variants of bs_kernel_fast_pf (FADB_BS_VARIANT=40 / 41, never the default): the same grid, trip structure and
// 13 streams with the arithmetic removed (MODE 1: what the memory system alone takes for this access pattern) and the same
// arithmetic on cache-resident operands with the stores predicated off (MODE 2: what the SMs alone take).  Their two
// times bracket the real kernel: max(mem, compute) is its floor, the gap to it is imperfect overlap.
template <int MODE>
__global__ void __launch_bounds__(256, 3)
bs_kernel_diag(size_t n, const double* __restrict__ S, const double* __restrict__ K,
               const double* __restrict__ sigma, const double* __restrict__ tau,
               const double* __restrict__ r, BSOut out)
{
    __shared__ double s_tab[675];
    for (int i = threadIdx.x; i < 256; i += 256) s_tab[i] = fm::kErfcx32_d[i];
    if (threadIdx.x < 32) s_tab[256 + threadIdx.x] = fm::kExp2_32_d[threadIdx.x];
    for (int i = threadIdx.x; i < 387; i += 256) s_tab[288 + i] = fm::kLogT_d[i];
    __syncthreads();
    pdl_prologue();
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    for (size_t i = tid; i < n; i += nthreads) {
        const size_t j = MODE == 2 ? (i & 4095) : i;
        const double s = __ldg(S + j), k = __ldg(K + j), g = __ldg(sigma + j), t = __ldg(tau + j), rr = __ldg(r + j);
        if (MODE == 1) {
            BSRes o;
            o.call = s + k; o.put = s - k; o.cS = g; o.cSig = t; o.cR = rr; o.pS = g + t; o.pSig = t + rr; o.pR = s + rr;
            bs_store1<false>(out, i, o);
        } else {
            const BSRes o = bs_eval_fast<true>(s, k, g, t, rr, s_tab);
            if (o.call == -12345.678) bs_store1<false>(out, i, o);     // never true: keeps the evaluation alive
        }
    }
}

// memory-only with VW options per thread per trip (128- / 256-bit accesses): does the access WIDTH move the floor?
template <int VW>
__global__ void __launch_bounds__(256)
bs_kernel_memdiag(size_t n, const double* __restrict__ S, const double* __restrict__ K,
                  const double* __restrict__ sigma, const double* __restrict__ tau,
                  const double* __restrict__ r, BSOut out)
{
    pdl_prologue();
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    const double* in[5] = {S, K, sigma, tau, r};
    for (size_t i = tid * VW; i + VW <= n; i += nthreads * VW) {
        double v[5][VW];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            if (VW == 4) { const double4_t w = ldg256_stream(in[k] + i); for (int e = 0; e < VW; ++e) v[k][e] = w.v[e]; }
            else { const double2 w = __ldg(reinterpret_cast<const double2*>(in[k] + i)); v[k][0] = w.x; v[k][VW - 1] = w.y; }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            double4_t w;
#pragma unroll
            for (int e = 0; e < VW; ++e) w.v[e] = v[k % 5][e] + v[(k + 1) % 5][e];
            if (VW == 4) st256(out.p[k] + i, w);
            else *reinterpret_cast<double2*>(out.p[k] + i) = make_double2(w.v[0], w.v[VW - 1]);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Bulk-copy (TMA) variant: ONE persistent CTA per SM; every group of G warps runs its own pipeline.
// A tile is G*32*OPL consecutive options.  The group's leader thread streams the five input tiles of the group's
// next tiles into a group-private STAGES-deep shared-memory ring with cp.async.bulk (completion on a group-private
// mbarrier); the threads read their option from shared memory (conflict-free LDS.64), evaluate it in registers
// exactly like bs_kernel_fast_pf, stage the eight results in a group-private tile, and the leader sends them out
// with cp.async.bulk shared -> global (accumulating adjoints: cp.reduce.async.bulk .add.f64, the same IEEE add
// performed at L2).  No per-thread global address arithmetic, no registers held for prefetched operands, no
// CTA-wide barrier in the loop (G = 1: __syncwarp only; G > 1: one named barrier per group), and the load depth
// is set by the ring instead of by occupancy.  Needs 16-byte aligned arrays; the n % tile tail and small batches
// take the per-thread path.
// ---------------------------------------------------------------------------------------
template <int WARPS, int G, int STAGES, int OPL, int OB>
struct BsTmaSmem {
    static constexpr int GROUPS = WARPS / G, TILE = G * 32 * OPL;
    double in[GROUPS][STAGES][5][TILE];
    double out[GROUPS][OB][8][TILE];
    double tab[676];                       // erfcx pieces [0, 256) + 2^(j/32) [256, 288) + log table [288, 675)
    uint64_t full[GROUPS][STAGES];
};

template <int G>
__device__ __forceinline__ void bs_group_sync(int grp)
{
    if (G == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(G * 32) : "memory");
}

template <bool ACC, int WARPS, int G, int STAGES, int OPL, int OB>
__global__ void __launch_bounds__(WARPS * 32, 1)
bs_kernel_tma(size_t n, const double* __restrict__ S, const double* __restrict__ K,
              const double* __restrict__ sigma, const double* __restrict__ tau,
              const double* __restrict__ r, BSOut out)
{
    static_assert(WARPS % G == 0 && (G == 1 || WARPS / G <= 15), "bs_kernel_tma: one named barrier per group");
    using Smem = BsTmaSmem<WARPS, G, STAGES, OPL, OB>;
    constexpr int TILE = Smem::TILE, GROUPS = Smem::GROUPS;
    constexpr unsigned TB = TILE * 8;
    extern __shared__ __align__(128) unsigned char bs_tma_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(bs_tma_raw);
    // warp index through a shuffle: the compiler then knows it (and the tile arithmetic derived from it) is warp-uniform
    // and issues the bulk copies from uniform registers instead of a per-lane waterfall loop around each of them
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int grp = warp / G, lig = (int)threadIdx.x - grp * (G * 32);      // thread index inside the group
    const bool leader = lig == 0;
    for (int i = threadIdx.x; i < 256; i += WARPS * 32) sm.tab[i] = fm::kErfcx32_d[i];
    if (threadIdx.x < 32) sm.tab[256 + threadIdx.x] = fm::kExp2_32_d[threadIdx.x];
    for (int i = threadIdx.x; i < 387; i += WARPS * 32) sm.tab[288 + i] = fm::kLogT_d[i];
    if (leader) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&sm.full[grp][s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    pdl_prologue();      // no global-memory access before the previous grid on the stream has completed

    const unsigned ntiles = (unsigned)(n / TILE);
    const unsigned stride = GROUPS * gridDim.x;
    const unsigned first = (unsigned)grp * gridDim.x + blockIdx.x;      // a partial last round spreads over all SMs
    uint64_t* bar = sm.full[grp];
    auto issue = [&](unsigned tile, int s) {                            // leader only
        mbar_arrive_expect_tx(&bar[s], 5 * TB);
        const size_t off = (size_t)tile * TILE;
        bulk_g2s(sm.in[grp][s][0], S + off, TB, &bar[s]);
        bulk_g2s(sm.in[grp][s][1], K + off, TB, &bar[s]);
        bulk_g2s(sm.in[grp][s][2], sigma + off, TB, &bar[s]);
        bulk_g2s(sm.in[grp][s][3], tau + off, TB, &bar[s]);
        bulk_g2s(sm.in[grp][s][4], r + off, TB, &bar[s]);
    };
    if (leader) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            const unsigned t = first + (unsigned)s * stride;
            if (t < ntiles) issue(t, s);
        }
    }
    int s = 0, ob = 0;
    unsigned par = 0;
    for (unsigned tile = first; tile < ntiles; tile += stride) {
        mbar_wait(&bar[s], par);
#pragma unroll
        for (int o = 0; o < OPL; ++o) {
            const int e = o * (G * 32) + lig;
            const double vs = sm.in[grp][s][0][e], vk = sm.in[grp][s][1][e], vg = sm.in[grp][s][2][e],
                         vt = sm.in[grp][s][3][e], vr = sm.in[grp][s][4][e];
            if (o == OPL - 1) {          // the stage is in registers: refill it with the tile STAGES rounds ahead
                bs_group_sync<G>(grp);
                if (leader) {
                    const unsigned nt = tile + STAGES * stride;
                    if (nt < ntiles) issue(nt, s);
                }
            }
            const BSRes o8 = bs_eval_fast<true>(vs, vk, vg, vt, vr, sm.tab);
            if (o == 0) {                // the result tile sent OB rounds ago has left shared memory: its buffer is free
                if (leader) bulk_wait_read<OB - 1>();
                bs_group_sync<G>(grp);
            }
            double* po = &sm.out[grp][ob][0][e];
            po[0 * TILE] = o8.call; po[1 * TILE] = o8.put;
            po[2 * TILE] = o8.cS;   po[3 * TILE] = o8.cSig; po[4 * TILE] = o8.cR;
            po[5 * TILE] = o8.pS;   po[6 * TILE] = o8.pSig; po[7 * TILE] = o8.pR;
        }
        fence_async_smem();
        bs_group_sync<G>(grp);
        if (leader) {
            const size_t off = (size_t)tile * TILE;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (ACC && k >= 2) bulk_s2g_add_f64(out.p[k] + off, sm.out[grp][ob][k], TB);
                else bulk_s2g(out.p[k] + off, sm.out[grp][ob][k], TB);
            }
            bulk_commit();
        }
        if (OB > 1) ob ^= 1;
        if (++s == STAGES) { s = 0; par ^= 1; }
    }
    if (leader) bulk_wait<0>();
    // the n % TILE options behind the last full tile: per-thread path, the last CTA
    if (blockIdx.x == gridDim.x - 1) {
        for (size_t i = (size_t)ntiles * TILE + threadIdx.x; i < n; i += WARPS * 32) {
            const BSRes o8 = bs_eval_fast<true>(__ldg(S + i), __ldg(K + i), __ldg(sigma + i), __ldg(tau + i), __ldg(r + i), sm.tab);
            bs_store1<ACC>(out, i, o8);
        }
    }
}

template <bool ACC, int WARPS, int G, int STAGES, int OPL, int OB>
static int launch_bs_tma(Context& c, cudaStream_t st, size_t n, const double* S, const double* K, const double* sigma,
                         const double* tau, const double* r, const BSOut& o, int max_grid)
{
    using Smem = BsTmaSmem<WARPS, G, STAGES, OPL, OB>;
    auto kern = bs_kernel_tma<ACC, WARPS, G, STAGES, OPL, OB>;
    constexpr size_t smem = sizeof(Smem);
    static_assert(smem <= 227 * 1024, "bs_kernel_tma: shared-memory ring exceeds 227 KB");
    static const cudaError_t attr = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    FADB_CUDA(attr);
    const size_t ntiles = n / Smem::TILE;
    size_t grid = (ntiles + Smem::GROUPS - 1) / Smem::GROUPS;
    if (grid > (size_t)c.sms) grid = c.sms;
    if (max_grid > 0 && grid > (size_t)max_grid) grid = max_grid;
    if (grid < 1) grid = 1;
    FADB_CUDA(launch_pdl(kern, dim3((unsigned)grid), dim3(WARPS * 32), smem, st, n, S, K, sigma, tau, r, o));
    return FADB_OK;
}

// host_mapped: the arrays are page-locked HOST memory (fadb_black_scholes_host): the PCIe link, not the
// FP64 pipe, is the limit there and FEWER resident CTAs are faster -- measured per 2^20 options:
// 444 CTAs 1.62 ms, 296 1.57, 148 1.54, 74 1.52, 24-48 1.506, 16 1.53, 8 1.56 (FADB_BS_HOST_GRID)
int launch_black_scholes(Context& c, cudaStream_t st, size_t n, const double* S, const double* K,
                         const double* sigma, const double* tau, const double* r,
                         double* const out[8], unsigned flags, bool host_mapped = false, int pdl_mode = PDL_PLAIN)
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
    // Measured on B200, 2^20 options per launch (us per launch): libdevice math 29.2 (variant 5);
    // fastmath 27.7 (variant 6); fastmath + software-pipelined loads, 128 registers, 2 CTAs/SM
    // 25.5 (variant 20); the same kernel launched with programmatic dependent launch 23.5 (variant 22);
    // Phi from the 32-interval degree-7 erfcx table in shared memory (40 fewer FP64 instructions per
    // option, 110 registers) 22.6 (variant 21); the same at 3 CTAs/SM (78 registers, no spills) 21.4;
    // 4 CTAs/SM (64 registers, spills) 22.7 (variant 24); exp through a 32-entry 2^(j/32) table in shared
    // memory (11 instead of 20 FP64 instructions, three per option) 20.7; log through a 129-entry
    // {1/c, log c} table (no division, 15 instead of 27) 20.5 (default).  4 CTAs of 224 threads at 72 registers
    // (7.9 instead of 9.2 trips per thread at 2^20 options) measured 21.4: removed.
    // Two options per thread, tighter launch bounds (spills), atomic work-stealing and streaming (st.global.cs)
    // result stores (20.68 against 20.48 us, A/B twice) were measured slower and removed.  So was a branch-free loop body
    // (exp_t without its early return + an unconditional prefetch: 347 instead of 405 instructions per option, one
    // straight-line block, 21.3 us) and exp_t / log_t literals from the constant bank (326 instructions, 22.0 us; without
    // the branch-free part 20.9-23.4 us): the kernel is bound by dependent-issue latency, not by issue slots.
    // (function-local statics with an initialiser: thread-safe on first use, C++11 [stmt.dcl])
    static const int variant = [] { const char* e = getenv("FADB_BS_VARIANT"); return e ? atoi(e) : 23; }();
    static const int host_grid = [] { const char* e = getenv("FADB_BS_HOST_GRID"); return e ? atoi(e) : 32; }();
    static const long tma_min = [] { const char* e = getenv("FADB_BS_TMA_MIN"); return e ? atol(e) : 148l * 24 * 32; }();
    const bool ovw = flags & FADB_OVERWRITE_ADJ;
    // bulk-copy pipeline (variants 30-39): 16-byte aligned arrays, at least one tile per warp
    if (variant >= 30 && variant < 40 && vec_ok && n >= (size_t)tma_min && (n >> 36) == 0) {
        const int mg = host_mapped ? host_grid : 0;
        int rc;
#define FADB_BS_TMA(W, G_, ST, OPL_, OB_) (ovw ? launch_bs_tma<false, W, G_, ST, OPL_, OB_>(c, st, n, S, K, sigma, tau, r, o, mg) \
                                               : launch_bs_tma<true, W, G_, ST, OPL_, OB_>(c, st, n, S, K, sigma, tau, r, o, mg))
        switch (variant) {
        case 31: rc = FADB_BS_TMA(24, 1, 2, 2, 1); break;
        case 32: rc = FADB_BS_TMA(28, 1, 3, 1, 1); break;
        case 33: rc = FADB_BS_TMA(32, 1, 2, 1, 2); break;
        case 34: rc = FADB_BS_TMA(24, 2, 3, 1, 2); break;
        case 35: rc = FADB_BS_TMA(24, 4, 3, 1, 2); break;
        case 36: rc = FADB_BS_TMA(24, 8, 3, 1, 2); break;
        case 37: rc = FADB_BS_TMA(28, 4, 3, 1, 1); break;
        case 38: rc = FADB_BS_TMA(24, 2, 2, 2, 1); break;
        case 39: rc = FADB_BS_TMA(32, 4, 2, 1, 2); break;
        default: rc = FADB_BS_TMA(24, 1, 3, 1, 2); break;
        }
#undef FADB_BS_TMA
        if (rc) return rc;
        c.launches++;
        return FADB_OK;
    }
    static const int host_wide = [] { const char* e = getenv("FADB_BS_HOST_WIDE"); return e ? atoi(e) : 0; }();
    bool wide_ok = host_wide == 2 || host_wide == 4;
    if (host_wide == 4) {
        uintptr_t bits = (uintptr_t)S | (uintptr_t)K | (uintptr_t)sigma | (uintptr_t)tau | (uintptr_t)r;
        for (int k = 0; k < 8; ++k) bits |= (uintptr_t)out[k];
        wide_ok = (bits & 31u) == 0;
    }
    if (host_mapped && ovw && vec_ok && wide_ok) {
        if (host_wide == 4) bs_kernel_host_wide<4><<<host_grid, 256, 0, st>>>(n, S, K, sigma, tau, r, o);
        else bs_kernel_host_wide<2><<<host_grid, 256, 0, st>>>(n, S, K, sigma, tau, r, o);
        c.launches++;
        FADB_CUDA(cudaGetLastError());
        return FADB_OK;
    }
#define FADB_BS_LAUNCH(KERNEL)                                                                          \
    do {                                                                                                \
        if (ovw) KERNEL<false><<<resident_grid(c, KERNEL<false>, threads, 0, n, threads), threads, 0, st>>>(n, S, K, sigma, tau, r, o); \
        else KERNEL<true><<<resident_grid(c, KERNEL<true>, threads, 0, n, threads), threads, 0, st>>>(n, S, K, sigma, tau, r, o);       \
    } while (0)
    if (variant == 43 || variant == 44) {       // memory-only, 2 / 4 options per thread per access
        if (variant == 43) FADB_CUDA(launch_pdl(bs_kernel_memdiag<2>, resident_grid(c, bs_kernel_memdiag<2>, threads, 0, n / 2, threads), threads, 0, st, n, S, K, sigma, tau, r, o));
        else FADB_CUDA(launch_pdl(bs_kernel_memdiag<4>, resident_grid(c, bs_kernel_memdiag<4>, threads, 0, n / 4, threads), threads, 0, st, n, S, K, sigma, tau, r, o));
        c.launches++;
        return FADB_OK;
    }
    if (variant == 42) {       // memory-only at the real kernel's grid (3 CTAs/SM)
        FADB_CUDA(launch_pdl(bs_kernel_diag<1>, std::min(3 * c.sms, (int)((n + threads - 1) / threads)), threads, 0, st, n, S, K, sigma, tau, r, o));
        c.launches++;
        return FADB_OK;
    }
    if (variant == 40 || variant == 41) {
        if (variant == 40) FADB_CUDA(launch_pdl(bs_kernel_diag<1>, resident_grid(c, bs_kernel_diag<1>, threads, 0, n, threads), threads, 0, st, n, S, K, sigma, tau, r, o));
        else FADB_CUDA(launch_pdl(bs_kernel_diag<2>, resident_grid(c, bs_kernel_diag<2>, threads, 0, n, threads), threads, 0, st, n, S, K, sigma, tau, r, o));
        c.launches++;
        return FADB_OK;
    }
    if (variant == 5) FADB_BS_LAUNCH(bs_kernel_libdevice);
    else if (variant == 6) FADB_BS_LAUNCH(bs_kernel_fast);
    else if (variant == 20) FADB_BS_LAUNCH(bs_kernel_fast_pf);
    else {
        // default: same kernel, launched with programmatic stream serialization so that back-to-back
        // batches overlap launch latency and the tail of the previous grid (the kernel waits on
        // cudaGridDependencySynchronize before its first memory access)
        if (variant == 22) {      // single degree-27 erfcx polynomial (23.5 us)
            if (ovw)
                FADB_CUDA(launch_pdl(bs_kernel_fast_pf<false>, resident_grid(c, bs_kernel_fast_pf<false>, threads, 0, n, threads),
                                     threads, 0, st, n, S, K, sigma, tau, r, o));
            else
                FADB_CUDA(launch_pdl(bs_kernel_fast_pf<true>, resident_grid(c, bs_kernel_fast_pf<true>, threads, 0, n, threads),
                                     threads, 0, st, n, S, K, sigma, tau, r, o));
#define FADB_BS_PW(MINB)                                                                                                 \
    do {                                                                                                                 \
        int grid_ = resident_grid(c, bs_kernel_fast_pf<false, true, MINB>, threads, 0, n, threads);                      \
        if (host_mapped && grid_ > host_grid) grid_ = host_grid;                                                         \
        if (ovw)                                                                                                         \
            FADB_CUDA(launch_pdl(bs_kernel_fast_pf<false, true, MINB>, grid_, threads, 0, st, n, S, K, sigma, tau, r, o)); \
        else                                                                                                             \
            FADB_CUDA(launch_pdl(bs_kernel_fast_pf<true, true, MINB>, grid_, threads, 0, st, n, S, K, sigma, tau, r, o)); \
    } while (0)
        } else if (variant == 21) { FADB_BS_PW(2);
        } else if (variant == 24) { FADB_BS_PW(4);
        } else if (pdl_mode == PDL_EARLY && ovw && !host_mapped && pdl_enabled()) {
            auto kern = bs_kernel_fast_pf<false, true, 3, 256, PDL_EARLY>;
            FADB_CUDA(launch_pdl(kern, resident_grid(c, kern, threads, 0, n, threads), threads, 0, st, n, S, K, sigma, tau, r, o));
        } else if (pdl_mode == PDL_HEAD && ovw && !host_mapped && pdl_enabled()) {
            auto kern = bs_kernel_fast_pf<false, true, 3, 256, PDL_HEAD>;
            FADB_CUDA(launch_pdl(kern, resident_grid(c, kern, threads, 0, n, threads), threads, 0, st, n, S, K, sigma, tau, r, o));
        } else { FADB_BS_PW(3);
        }
    }
#undef FADB_BS_LAUNCH
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

// `count` batches in ONE host call (the per-option loop of example/black_scholes.cpp:43-70 over count x n options): batch b
// reads the operand set (first_set + b) % nsets -- in[5 * s + {0..4}] = S, K, sigma, tau, r of set s, out[8 * s + k] its
// eight result arrays -- with one kernel launch per batch, back to back on the library stream, no host work in between.
extern "C" int fadb_black_scholes_batches(size_t count, size_t n, size_t nsets, size_t first_set,
                                          const double* const* in, double* const* out, unsigned flags)
{
    FADB_REQUIRE(nsets >= 1 && in != nullptr && out != nullptr, "fadb_black_scholes_batches: null operand tables");
    if (n > 0)
        for (size_t s = 0; s < nsets; ++s) {
            for (int k = 0; k < 5; ++k) FADB_REQUIRE(in[5 * s + k] != nullptr, "fadb_black_scholes_batches: null input");
            for (int k = 0; k < 8; ++k) FADB_REQUIRE(out[8 * s + k] != nullptr, "fadb_black_scholes_batches: null output");
        }
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    // Batches 1.. start without waiting for the previous batch's grid (see bs_kernel_fast_pf<EARLY>): allowed when consecutive
    // batches work on DIFFERENT operand sets (nsets >= 2: the caller's sets must not alias each other) and results are
    // stored, not accumulated.  FADB_BS_EARLY=0 keeps every batch on the ordinary wait-first kernel.
    static const bool early_ok = [] { const char* e = getenv("FADB_BS_EARLY"); return !(e && e[0] == '0'); }();
    const bool can_early = early_ok && nsets >= 2 && (flags & FADB_OVERWRITE_ADJ);
    for (size_t b = 0; b < count; ++b) {
        const size_t s = (first_set + b) % nsets;
        rc = launch_black_scholes(*c, c->stream, n, in[5 * s], in[5 * s + 1], in[5 * s + 2], in[5 * s + 3], in[5 * s + 4],
                                  out + 8 * s, flags, false, !can_early || count < 2 ? PDL_PLAIN : (b > 0 ? PDL_EARLY : PDL_HEAD));
        if (rc) return rc;
    }
    return FADB_OK;
}

struct Tuner { size_t n = 0; int calls = 0; double best[2] = {1e30, 1e30}; int choice = -1; };
static Tuner g_bs_host_tuner[kMaxDevices];

// Which transfer scheme fadb_black_scholes_host currently uses on this device: 1 = one launch over PCIe on the page-locked
// arrays, 0 = staged through the copy engines, -1 = still measuring (first six calls); out_ms[2] = best call time seen per scheme.
extern "C" int fadb_black_scholes_host_mode(double out_ms[2])
{
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    const Tuner& tu = g_bs_host_tuner[c->device];
    if (out_ms) { out_ms[0] = tu.best[0] < 1e29 ? tu.best[0] * 1e3 : -1.0; out_ms[1] = tu.best[1] < 1e29 ? tu.best[1] * 1e3 : -1.0; }
    return tu.choice;
}

// Host-buffer entry: the call a user of the reference makes with plain host arrays.
// Page-locked arrays: one launch of the resident kernel over PCIe.  Pageable arrays: the batch is
// chunked and H2D / kernel / D2H overlap on three streams' worth of double buffering.
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
    static const int chunk_log2 = [] { const char* e = getenv("FADB_BS_HOST_CHUNK_LOG2"); return e ? atoi(e) : 18; }();
    const size_t cmax = (size_t)1 << chunk_log2;
    const size_t chunk = n < cmax ? n : cmax;
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
    // Pageable host memory is accepted: cudaMemcpyAsync stages it, chunk by chunk (below).
    // When all 13 arrays are page-locked (cudaHostAlloc / cudaHostRegister, what torch's pin_memory()
    // gives) the resident kernel runs directly on their UVA device aliases instead: ONE launch reads
    // the inputs and writes the results over PCIe, both directions busy at once, no staging buffers and
    // no copy calls.  Measured on B200, 2^20 options (104 MB over PCIe): staged 1.78 ms at the best
    // chunk size (4 chunks; more chunks cost ~35 us each in copy calls), direct 1.51 ms (32 CTAs); plain
    // cudaMemcpy of the 64 MB of results alone takes 1.18 ms.  A mix (kernel reads host memory, DMA
    // write-back) measured slower than both (2.04 ms), and so did the opposite mix (copy engine brings the inputs in
    // chunk by chunk, the kernel stores straight to the host arrays): 1.72 / 1.87 / 2.17 ms at 4 / 8 / 16 chunks,
    // i.e. ~37 us per chunk for its five small serial copies, whatever the CTA count.  FADB_BS_HOST_MODE=0 forces
    // the staged path.
    const double* hin[5] = {hS, hK, hsigma, htau, hr};
    // FADB_BS_HOST_MODE: 0 = always staged (copy engines), 1 = always direct (one launch over PCIe) when the arrays are
    // page-locked, 2 (default) = measure both on the first calls and keep the faster.  Which one wins depends on what else
    // loads the host: alone on the box the direct launch is ~15 % faster (1.51 vs 1.78 ms per 2^20 options); with 8 ranks
    // pulling through one root complex the copy engines' large transfers can win.  The tuner times calls 2..5 (two of
    // each mode, after one warm-up call of each) with the host clock -- the call is synchronous, so that is its full
    // cost -- and re-tunes when the batch size changes by more than 2x.  Both modes return identical bits.
    static const int host_mode = [] { const char* e = getenv("FADB_BS_HOST_MODE"); return e ? atoi(e) : 2; }();
    Tuner& tu = g_bs_host_tuner[c->device];
    int want = host_mode == 0 ? 0 : 1;                  // 1 = direct
    bool timing = false;
    if (host_mode == 2 && n >= ((size_t)1 << 18)) {
        if (tu.n == 0 || n > 2 * tu.n || 2 * n < tu.n) { tu = Tuner(); tu.n = n; }
        if (tu.choice < 0) {
            want = (tu.calls & 1) ? 0 : 1;              // direct, staged, direct, staged, ...
            timing = tu.calls >= 2;
        } else want = tu.choice;
    }
    const auto t_begin = std::chrono::steady_clock::now();
    auto tuner_done = [&](int mode_used) {
        if (host_mode != 2 || tu.choice >= 0 || n < ((size_t)1 << 18)) return;
        if (timing) {
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
            if (dt < tu.best[mode_used]) tu.best[mode_used] = dt;
        }
        if (++tu.calls >= 6) tu.choice = tu.best[1] <= tu.best[0] ? 1 : 0;
    };
    const double* din[5];
    double* dout_zc[8];
    bool locked = want != 0;
    for (int k = 0; k < 13 && locked; ++k) {
        cudaPointerAttributes at;
        const void* hp = k < 5 ? (const void*)hin[k] : (const void*)hout[k - 5];
        if (cudaPointerGetAttributes(&at, hp) != cudaSuccess) { cudaGetLastError(); locked = false; break; }
        if (at.type != cudaMemoryTypeHost || !at.devicePointer) { locked = false; break; }
        if (k < 5) din[k] = static_cast<const double*>(at.devicePointer);
        else dout_zc[k - 5] = static_cast<double*>(at.devicePointer);
    }
    if (locked) {
        rc = launch_black_scholes(*c, c->stream, n, din[0], din[1], din[2], din[3], din[4], dout_zc, FADB_OVERWRITE_ADJ, true);
        if (rc) return rc;
        FADB_CUDA(cudaStreamSynchronize(c->stream));
        tuner_done(1);
        return FADB_OK;
    }
    if (want != 0 && host_mode == 2 && tu.choice < 0) tu.choice = 0;     // pageable arrays: nothing to choose
    if (dcap < chunk) {
        for (int b = 0; b < NBUF; ++b) {
            if (dbuf[b]) FADB_CUDA(cudaFree(dbuf[b]));
            FADB_CUDA(cudaMalloc(&dbuf[b], sizeof(double) * 13 * chunk));
        }
        dcap = chunk;
    }
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
    tuner_done(0);
    return FADB_OK;
}
