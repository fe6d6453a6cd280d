// normal.cu — K4: ad::normal_adj_log_pdf(x, mu, sigma) value + adjoints, leaf operands.
//
// Replaces NormalAdjLogPDFNode<..., tuple<scl|vec, scl|vec, scl|vec>>::feval/beval
// (reverse/stat/normal.hpp: sss :108-180, vss :182-301, vvs :304-381, vsv :384-495,
// vvv :498-578) and the VarView::beval `adj += seed` of its operands (var_view.hpp:91-94).
//
// The reference makes 2 passes over sigma (validity + log-sum), one over x,mu,sigma for the
// value and three adjoint passes that each re-read x,mu,sigma (~160 B/element for vvv).
// Here the root-level seed is known when the call is made, so value and all adjoints are
// produced by ONE fused pass: 3 streaming reads + 3 read-modify-writes = 72 B/element (vvv).
// A vector sigma must be valid EVERYWHERE for any adjoint to be written (normal.hpp:440-442,457:
// sigma_i <= 0 => value -inf and NO adjoint update).  The pass is therefore OPTIMISTIC: it writes the
// adjoints of every element whose own sigma_i is valid and counts the invalid ones in the same sweep; when the
// (global) count turns out non-zero -- the rare error case -- normal_undo_kernel takes the contributions back
// (recomputed bit-identically, subtracted).  The common case stays at 72 B/element; the undo launch exits on its
// first load otherwise.  With zero-initialised or FADB_OVERWRITE_ADJ adjoints the undo is exact (they read 0
// afterwards); accumulated into non-zero adjoints it restores them to within one rounding of (previous +
// contribution).  FADB_STRICT_SIGMA buys the bit-exact "untouched" with a read-only pre-pass over sigma
// (+8 B/element, sigma_check_kernel); FADB_TRUST_SIGMA skips the undo launch altogether.
#include "common.cuh"
#include "p2p_device.cuh"

namespace fadb {

constexpr int NCH = 4;  // channels: 0 = sum d^2 | z^2, 1 = sum log sigma_i, 2 = sum d | d/s^2, 3 = #invalid

struct NormalArgs {
    size_t n;
    const double* x; const double* mu; const double* sigma;
    double* x_adj; double* mu_adj; double* sigma_adj;
    double seed;
    int overwrite;
    const double* invalid_in;   // device count of sigma_i <= 0 from the pre-pass (may be null)
};

// ------------------------------------------------------------------ sigma validity pre-pass
__global__ void __launch_bounds__(512)
sigma_check_kernel(size_t n, const double* __restrict__ sigma, double* partials,
                   unsigned int* ticket, double* out_count)
{
    pdl_prologue();
    __shared__ double smem[32];
    double acc[1] = {0.0};
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t nth = (size_t)gridDim.x * blockDim.x;
    const bool vec = aligned32(sigma);
    const size_t n4 = vec ? n / 4 : 0;
    size_t q = tid;
    for (; q + 3 * nth < n4; q += 4 * nth) {          // read-only stream: four 256-bit loads in flight per thread
        double4_t s[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) s[u] = ldg256_stream(sigma + 4 * (q + u * nth));
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[0] += (s[u].v[k] > 0) ? 0.0 : 1.0;
    }
    for (; q < n4; q += nth) {
        const double4_t s = ldg256_stream(sigma + 4 * q);
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[0] += (s.v[k] > 0) ? 0.0 : 1.0;
    }
    for (size_t i = n4 * 4 + tid; i < n; i += nth) acc[0] += (sigma[i] > 0) ? 0.0 : 1.0;
    block_sum<1>(acc, smem);
    grid_finish<1>(acc, smem, partials, ticket, [&](double (&tot)[1]) { out_count[0] = tot[0]; });
}

// ------------------------------------------------------------------ fused main pass
template <bool MU_VEC, bool SIG_VEC>
struct Elem {
    // per-element forward terms + adjoints, formulas in the order the reference writes them
    __device__ __forceinline__ static void run(double x, double m, double s, double seed,
                                               double inv_s_sq /* scalar sigma only */,
                                               double (&acc)[NCH], double& gx, double& gm, double& gs)
    {
        const double d = x - m;
        if (SIG_VEC) {
            const double z = d / s;                           // normal.hpp:449, 546
            acc[0] += z * z;
            acc[1] += log(s);                                 // normal.hpp:483, 572
            const double s2 = s * s;
            if (MU_VEC) {
                gx = seed * (m - x) / s2;                     // vvv, normal.hpp:564
                gm = seed * d / s2;                           // vvv, normal.hpp:563
            } else {
                gx = (seed / s2) * (m - x);                   // vsv, normal.hpp:474
                acc[2] += d / s2; gm = 0.0;                   // vsv, normal.hpp:471
            }
            gs = (seed / s) * (z * z - 1.);                   // normal.hpp:468, 560
        } else {
            acc[0] += d * d;                                  // squaredNorm, normal.hpp:244, 351
            const double c = seed * inv_s_sq;
            gx = c * (m - x);                                 // normal.hpp:282, 371
            if (MU_VEC) gm = c * d;                           // normal.hpp:370
            else { acc[2] += d; gm = 0.0; }                   // normal.hpp:277
            gs = 0.0;
        }
    }
};

__device__ __forceinline__ void rmw4(double* p, const double (&g)[4], int overwrite)
{
    double4_t v;
    if (overwrite) {
#pragma unroll
        for (int k = 0; k < 4; ++k) v.v[k] = g[k];
    } else {
        v = ld256(p);
#pragma unroll
        for (int k = 0; k < 4; ++k) v.v[k] += g[k];
    }
    st256(p, v);
}

template <bool MU_VEC, bool SIG_VEC>
__global__ void __launch_bounds__(256)
normal_kernel(NormalArgs a, double* partials, unsigned int* ticket, double* out_partial, P2PLink link)
{
    pdl_prologue();
    __shared__ double smem[32 * NCH];
    double acc[NCH] = {0.0, 0.0, 0.0, 0.0};
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t nth = (size_t)gridDim.x * blockDim.x;

    // Validity (uniform across the grid): scalar sigma is read by everyone; vector sigma uses
    // the pre-pass count.  Invalid => forward sums still flow (finish() turns them into -inf)
    // but no adjoint is touched.
    double s_scl = 1.0, m_scl = 0.0, inv_s_sq = 1.0;
    bool valid = true;
    if (!SIG_VEC) {
        s_scl = __ldg(a.sigma);
        valid = s_scl > 0;
        const double inv_s = 1. / s_scl;                      // normal.hpp:254-255
        inv_s_sq = inv_s * inv_s;
    } else if (a.invalid_in) {
        valid = __ldcg(a.invalid_in) == 0.0;
    }
    if (!MU_VEC) m_scl = __ldg(a.mu);
    const bool back = valid && a.seed != 0.0;                  // normal.hpp:160, 252, 457
    const bool wx = back && a.x_adj, wm = back && MU_VEC && a.mu_adj,
               ws = back && SIG_VEC && a.sigma_adj;

    bool vec = aligned32(a.x) && (!MU_VEC || aligned32(a.mu)) && (!SIG_VEC || aligned32(a.sigma)) &&
               (!a.x_adj || aligned32(a.x_adj)) && (!MU_VEC || !a.mu_adj || aligned32(a.mu_adj)) &&
               (!SIG_VEC || !a.sigma_adj || aligned32(a.sigma_adj));
    const size_t n4 = vec ? a.n / 4 : 0;
    size_t q0 = tid;
    if (!MU_VEC && !SIG_VEC && !wx) {
        // constant x, scalar mu and sigma: a pure reduction (8 B/elem).  Four independent 256-bit loads per
        // thread and trip keep the read stream deep; same per-thread accumulation order as the loop below.
        for (; q0 + 3 * nth < n4; q0 += 4 * nth) {
            double4_t xv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) xv[u] = ldg256_stream(a.x + 4 * (q0 + u * nth));
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double d = xv[u].v[k] - m_scl;
                    acc[0] += d * d;
                    acc[2] += d;
                }
        }
    }
    for (size_t q = q0; q < n4; q += nth) {
        const size_t i = 4 * q;
        const double4_t xv = ldg256_stream(a.x + i);
        double4_t mv, sv;
        if (MU_VEC) mv = ldg256_stream(a.mu + i);
        if (SIG_VEC) sv = ldg256_stream(a.sigma + i);
        double gx[4], gm[4], gs[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double s = SIG_VEC ? sv.v[k] : s_scl;
            Elem<MU_VEC, SIG_VEC>::run(xv.v[k], MU_VEC ? mv.v[k] : m_scl, s, a.seed, inv_s_sq, acc,
                                       gx[k], gm[k], gs[k]);
            if (SIG_VEC && !(s > 0)) { acc[3] += 1.0; gx[k] = gm[k] = gs[k] = 0.0; }   // optimistic pass: an invalid element adds nothing
        }
        if (wx) rmw4(a.x_adj + i, gx, a.overwrite);
        if (wm) rmw4(a.mu_adj + i, gm, a.overwrite);
        if (ws) rmw4(a.sigma_adj + i, gs, a.overwrite);
    }
    for (size_t i = n4 * 4 + tid; i < a.n; i += nth) {
        const double s = SIG_VEC ? a.sigma[i] : s_scl;
        double gx, gm, gs;
        Elem<MU_VEC, SIG_VEC>::run(a.x[i], MU_VEC ? a.mu[i] : m_scl, s, a.seed, inv_s_sq, acc, gx, gm, gs);
        if (SIG_VEC && !(s > 0)) { acc[3] += 1.0; gx = gm = gs = 0.0; }
        if (wx) a.x_adj[i] = a.overwrite ? gx : a.x_adj[i] + gx;
        if (wm) a.mu_adj[i] = a.overwrite ? gm : a.mu_adj[i] + gm;
        if (ws) a.sigma_adj[i] = a.overwrite ? gs : a.sigma_adj[i] + gs;
    }
    if (!SIG_VEC && !valid && blockIdx.x == 0 && threadIdx.x == 0) acc[3] += 1.0;
    block_sum<NCH>(acc, smem);
    grid_finish<NCH>(acc, smem, partials, ticket, [&](double (&tot)[NCH]) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) out_partial[c] = tot[c];
        // fused exchange: the partial sums go straight into every rank's mailbox over NVLink (p2p_device.cuh)
        if (link.world) p2p_push1(link, tot, NCH);
    });
}

// ------------------------------------------------------------------ undo of the optimistic pass (vector sigma)
// Runs behind the finish kernel; `count` = global number of sigma_i <= 0.  Zero (the common case): every thread
// leaves after one load.  Otherwise the per-element contributions are recomputed exactly as normal_kernel formed
// them (same Elem::run, same flags) and taken back: stored adjoints are zeroed, accumulated ones get them subtracted.
__device__ __forceinline__ void undo4(double* p, const double (&g)[4], int overwrite)
{
    double4_t v;
    if (overwrite) {
#pragma unroll
        for (int k = 0; k < 4; ++k) v.v[k] = 0.0;
    } else {
        v = ld256(p);
#pragma unroll
        for (int k = 0; k < 4; ++k) v.v[k] -= g[k];
    }
    st256(p, v);
}

template <bool MU_VEC>
__global__ void __launch_bounds__(256)
normal_undo_kernel(NormalArgs a, const double* count)
{
    pdl_prologue();
    if (__ldcg(count) == 0.0 || a.seed == 0.0) return;
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t nth = (size_t)gridDim.x * blockDim.x;
    const double m_scl = MU_VEC ? 0.0 : __ldg(a.mu);
    const bool wx = a.x_adj, wm = MU_VEC && a.mu_adj, ws = a.sigma_adj;
    const bool vec = aligned32(a.x) && (!MU_VEC || aligned32(a.mu)) && aligned32(a.sigma) &&
                     (!a.x_adj || aligned32(a.x_adj)) && (!MU_VEC || !a.mu_adj || aligned32(a.mu_adj)) &&
                     (!a.sigma_adj || aligned32(a.sigma_adj));
    const size_t n4 = vec ? a.n / 4 : 0;
    double unused[NCH] = {0.0, 0.0, 0.0, 0.0};
    for (size_t q = tid; q < n4; q += nth) {
        const size_t i = 4 * q;
        const double4_t xv = ldg256_stream(a.x + i), sv = ldg256_stream(a.sigma + i);
        double4_t mv;
        if (MU_VEC) mv = ldg256_stream(a.mu + i);
        double gx[4], gm[4], gs[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            Elem<MU_VEC, true>::run(xv.v[k], MU_VEC ? mv.v[k] : m_scl, sv.v[k], a.seed, 1.0, unused, gx[k], gm[k], gs[k]);
            if (!(sv.v[k] > 0)) gx[k] = gm[k] = gs[k] = 0.0;
        }
        if (wx) undo4(a.x_adj + i, gx, a.overwrite);
        if (wm) undo4(a.mu_adj + i, gm, a.overwrite);
        if (ws) undo4(a.sigma_adj + i, gs, a.overwrite);
    }
    for (size_t i = n4 * 4 + tid; i < a.n; i += nth) {
        const double s = a.sigma[i];
        double gx, gm, gs;
        Elem<MU_VEC, true>::run(a.x[i], MU_VEC ? a.mu[i] : m_scl, s, a.seed, 1.0, unused, gx, gm, gs);
        if (!(s > 0)) gx = gm = gs = 0.0;
        if (wx) a.x_adj[i] = a.overwrite ? 0.0 : a.x_adj[i] - gx;
        if (wm) a.mu_adj[i] = a.overwrite ? 0.0 : a.mu_adj[i] - gm;
        if (ws) a.sigma_adj[i] = a.overwrite ? 0.0 : a.sigma_adj[i] - gs;
    }
}

// ------------------------------------------------------------------ scalar epilogue
// Turns the (possibly all-reduced) partial sums into the node value and the adjoints of
// scalar operands.  n_total is the global element count.
__global__ void normal_finish_kernel(int shape_code, double n_total, const double* partial,
                                     const double* x, const double* mu, const double* sigma,
                                     double* x_adj, double* mu_adj, double* sigma_adj, double seed,
                                     int overwrite, double* out_value, double* out_invalid, P2PLink link)
{
    double pulled[NCH];
    if (link.world) {       // fused exchange: wait for every rank's push, reduce in rank order
        p2p_pull1(link, pulled, NCH);
        partial = pulled;
    }
    out_invalid[0] = partial[3];      // global #sigma<=0: what normal_undo_kernel behind this launch looks at
    const bool mu_vec = shape_code & 1, sig_vec = shape_code & 2;
    // normal.hpp:147, 440-442; a NaN count is the poison of a failed exchange (p2p_device.cuh) and stays NaN
    if (partial[3] != 0.0) { out_value[0] = (partial[3] != partial[3]) ? partial[3] : -INFINITY; return; }
    if (n_total == 1.0 && shape_code == 0) {
        // sss, normal.hpp:141-171
        const double s = sigma[0], inv_s = 1. / s;
        const double z = (x[0] - mu[0]) / s;
        out_value[0] = -0.5 * z * z - log(s);
        if (seed == 0.0) return;
        const double zb = (x[0] - mu[0]) * inv_s;
        if (sigma_adj) { const double g = seed * (zb * zb - 1) * inv_s; sigma_adj[0] = overwrite ? g : sigma_adj[0] + g; }
        const double adj = seed * zb * inv_s;
        if (mu_adj) mu_adj[0] = overwrite ? adj : mu_adj[0] + adj;
        if (x_adj) x_adj[0] = overwrite ? -adj : x_adj[0] - adj;
        return;
    }
    if (sig_vec) {
        out_value[0] = -0.5 * partial[0] - partial[1];              // normal.hpp:451, 548
        if (!mu_vec && mu_adj && seed != 0.0) {
            const double g = seed * partial[2];                    // normal.hpp:471-472
            mu_adj[0] = overwrite ? g : mu_adj[0] + g;
        }
    } else {
        const double s = sigma[0];
        const double inv_s = 1. / s, inv_s_sq = inv_s * inv_s;
        const double z_sq = partial[0] / (s * s);                   // normal.hpp:245, 351
        out_value[0] = -0.5 * z_sq - n_total * log(s);
        if (seed == 0.0) return;
        if (sigma_adj) {
            const double g = seed * (z_sq - n_total) * inv_s;      // normal.hpp:273, 365
            sigma_adj[0] = overwrite ? g : sigma_adj[0] + g;
        }
        if (!mu_vec && mu_adj) {
            const double g = seed * (partial[2] * inv_s_sq);       // normal.hpp:277-278
            mu_adj[0] = overwrite ? g : mu_adj[0] + g;
        }
    }
}

static int launch_partial(Context& c, int shape_code, const NormalArgs& a, double* out_partial, const P2PLink* lk = nullptr)
{
    P2PLink link;
    link.world = 0;
    if (lk) link = *lk;
    const int threads = 256;
    const size_t work = (a.n + 3) / 4;
    double* part = c.partials;
    unsigned int* ticket = c.tickets + 1;
#define FADB_LAUNCH_NORMAL(MV, SV)                                                               \
    FADB_CUDA(launch_pdl(normal_kernel<MV, SV>, resident_grid(c, normal_kernel<MV, SV>, threads, 0, work, threads), \
                         threads, 0, c.stream, a, part, ticket, out_partial, link))
    switch (shape_code & 3) {
    case 0: FADB_LAUNCH_NORMAL(false, false); break;
    case 1: FADB_LAUNCH_NORMAL(true, false); break;
    case 2: FADB_LAUNCH_NORMAL(false, true); break;
    case 3: FADB_LAUNCH_NORMAL(true, true); break;
    }
#undef FADB_LAUNCH_NORMAL
    c.launches++;
    FADB_CUDA(cudaGetLastError());
    return FADB_OK;
}

}  // namespace fadb

using namespace fadb;

static int check_normal_args(int shape_code, size_t n, const double* x, const double* mu,
                             const double* sigma)
{
    FADB_REQUIRE(shape_code >= 0 && shape_code <= 3, "normal_lpdf: shape_code must be 0..3");
    FADB_REQUIRE(n == 0 || (x && mu && sigma), "normal_lpdf: null operand");
    // scalar operands are dereferenced by the kernels whatever n is (an empty x still has a mu and a sigma)
    FADB_REQUIRE(((shape_code & 1) || mu) && ((shape_code & 2) || sigma), "normal_lpdf: null scalar operand");
    return FADB_OK;
}

extern "C" int fadb_sigma_check(size_t n, const double* sigma, double* dev_count)
{
    FADB_REQUIRE(dev_count && (n == 0 || sigma), "fadb_sigma_check: null pointer");
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    const int threads = 512;
    const int grid = resident_grid(*c, sigma_check_kernel, threads, 0, (n + 3) / 4, threads);
    FADB_CUDA(launch_pdl(sigma_check_kernel, grid, threads, 0, c->stream, n, sigma, c->partials, c->tickets + 0, dev_count));
    c->launches++;
    FADB_CUDA(cudaGetLastError());
    return FADB_OK;
}

extern "C" int fadb_normal_lpdf_partial(int shape_code, size_t n, const double* x,
                                        const double* mu, const double* sigma, double* x_adj,
                                        double* mu_adj, double* sigma_adj, double seed,
                                        unsigned flags, const double* dev_invalid_in,
                                        double* dev_partial4)
{
    int rc = check_normal_args(shape_code, n, x, mu, sigma);
    if (rc) return rc;
    FADB_REQUIRE(dev_partial4 != nullptr, "normal_lpdf_partial: null partial buffer");
    Context* c;
    rc = current_context(&c);
    if (rc) return rc;
    NormalArgs a{n, x, mu, sigma, x_adj, mu_adj, sigma_adj, seed, (flags & FADB_OVERWRITE_ADJ) ? 1 : 0,
                 dev_invalid_in};
    return launch_partial(*c, shape_code, a, dev_partial4);
}

// The same fused pass with the exchange fused in: the CTA that forms the four partial sums stores them straight
// into every rank's mailbox over NVLink (csrc/p2p.cu).  Pair with fadb_normal_lpdf_finish_pull.
extern "C" int fadb_normal_lpdf_partial_push(int shape_code, size_t n, const double* x, const double* mu, const double* sigma,
                                             double* x_adj, double* mu_adj, double* sigma_adj, double seed, unsigned flags,
                                             const double* dev_invalid_in, double* dev_partial4)
{
    int rc = check_normal_args(shape_code, n, x, mu, sigma);
    if (rc) return rc;
    FADB_REQUIRE(dev_partial4 != nullptr, "normal_lpdf_partial_push: null partial buffer");
    Context* c;
    rc = current_context(&c);
    if (rc) return rc;
    P2PLink link;
    if ((rc = p2p_next_link(*c, &link))) return rc;
    NormalArgs a{n, x, mu, sigma, x_adj, mu_adj, sigma_adj, seed, (flags & FADB_OVERWRITE_ADJ) ? 1 : 0,
                 dev_invalid_in};
    return launch_partial(*c, shape_code, a, dev_partial4, &link);
}

static int normal_finish_impl(int shape_code, size_t n_total, const double* dev_partial4, const double* x, const double* mu,
                              const double* sigma, double* x_adj, double* mu_adj, double* sigma_adj, double seed, unsigned flags,
                              double* host_value, bool pull)
{
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    P2PLink link;
    link.world = 0;
    if (pull && (rc = p2p_current_link(*c, &link))) return rc;
    double* dval = c->scalars + 0;
    normal_finish_kernel<<<1, 1, 0, c->stream>>>(shape_code, (double)n_total, dev_partial4, x, mu, sigma,
                                                 x_adj, mu_adj, sigma_adj, seed,
                                                 (flags & FADB_OVERWRITE_ADJ) ? 1 : 0, dval, c->scalars + 12, link);
    c->launches++;
    FADB_CUDA(cudaGetLastError());
    if (host_value) {
        FADB_CUDA(cudaMemcpyAsync(c->host_scalars, dval, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        FADB_CUDA(cudaStreamSynchronize(c->stream));
        *host_value = c->host_scalars[0];
    }
    return FADB_OK;
}

extern "C" int fadb_normal_lpdf_finish(int shape_code, size_t n_total, const double* dev_partial4,
                                       const double* x, const double* mu, const double* sigma,
                                       double* x_adj, double* mu_adj, double* sigma_adj,
                                       double seed, unsigned flags, double* host_value)
{
    return normal_finish_impl(shape_code, n_total, dev_partial4, x, mu, sigma, x_adj, mu_adj, sigma_adj, seed, flags, host_value, false);
}

// Consumer of fadb_normal_lpdf_partial_push: waits for every rank's push, sums the mailbox slots in rank order
// and forms the value and the scalar adjoints in the one finish kernel.
extern "C" int fadb_normal_lpdf_finish_pull(int shape_code, size_t n_total, const double* x, const double* mu, const double* sigma,
                                            double* x_adj, double* mu_adj, double* sigma_adj, double seed, unsigned flags,
                                            double* host_value)
{
    return normal_finish_impl(shape_code, n_total, nullptr, x, mu, sigma, x_adj, mu_adj, sigma_adj, seed, flags, host_value, true);
}

// Takes back what an optimistic pass over a vector sigma wrote when the global count of sigma_i <= 0 is non-zero
// (dev_invalid_count; NULL = the count the last finish kernel on this device published).  No-op for a scalar sigma.
extern "C" int fadb_normal_lpdf_undo(int shape_code, size_t n, const double* x, const double* mu, const double* sigma,
                                     double* x_adj, double* mu_adj, double* sigma_adj, double seed, unsigned flags,
                                     const double* dev_invalid_count)
{
    int rc = check_normal_args(shape_code, n, x, mu, sigma);
    if (rc) return rc;
    if (!(shape_code & 2) || seed == 0.0 || n == 0) return FADB_OK;
    const bool mu_vec = shape_code & 1;
    if (!x_adj && !(mu_vec && mu_adj) && !sigma_adj) return FADB_OK;
    Context* c;
    rc = current_context(&c);
    if (rc) return rc;
    NormalArgs a{n, x, mu, sigma, x_adj, mu_vec ? mu_adj : nullptr, sigma_adj, seed, (flags & FADB_OVERWRITE_ADJ) ? 1 : 0, nullptr};
    const double* cnt = dev_invalid_count ? dev_invalid_count : c->scalars + 12;
    const int threads = 256;
    if (mu_vec)
        FADB_CUDA(launch_pdl(normal_undo_kernel<true>, resident_grid(*c, normal_undo_kernel<true>, threads, 0, (n + 3) / 4, threads),
                             threads, 0, c->stream, a, cnt));
    else
        FADB_CUDA(launch_pdl(normal_undo_kernel<false>, resident_grid(*c, normal_undo_kernel<false>, threads, 0, (n + 3) / 4, threads),
                             threads, 0, c->stream, a, cnt));
    c->launches++;
    FADB_CUDA(cudaGetLastError());
    return FADB_OK;
}

extern "C" int fadb_normal_lpdf(int shape_code, size_t n, const double* x, const double* mu,
                                const double* sigma, double* x_adj, double* mu_adj,
                                double* sigma_adj, double seed, unsigned flags, double* host_value)
{
    int rc = check_normal_args(shape_code, n, x, mu, sigma);
    if (rc) return rc;
    Context* c;
    rc = current_context(&c);
    if (rc) return rc;
    double* partial = c->scalars + 8;     // 4 doubles
    double* invalid = c->scalars + 12;    // 1 double
    const bool sig_vec = shape_code & 2, mu_vec = shape_code & 1;
    // vector adjoints are written by the pass itself: with a vector sigma they need the all-elements validity verdict
    const bool guarded = sig_vec && (x_adj || (mu_vec && mu_adj) || sigma_adj) && seed != 0.0 && !(flags & FADB_TRUST_SIGMA);
    const double* invalid_in = nullptr;
    if (guarded && (flags & FADB_STRICT_SIGMA)) {          // verdict first (read-only pre-pass), then the pass
        rc = fadb_sigma_check(n, sigma, invalid);
        if (rc) return rc;
        invalid_in = invalid;
    }
    NormalArgs a{n, x, mu, sigma, x_adj, mu_adj, sigma_adj, seed, (flags & FADB_OVERWRITE_ADJ) ? 1 : 0,
                 invalid_in};
    const bool sss = (n == 1 && shape_code == 0);
    if (sss) {
        // single element: the epilogue does everything (normal.hpp:141-171)
        FADB_CUDA(cudaMemsetAsync(partial, 0, 4 * sizeof(double), c->stream));
        // sigma <= 0 is detected in the epilogue through partial[3]; compute it on device
        a.x_adj = a.mu_adj = a.sigma_adj = nullptr;
        rc = launch_partial(*c, shape_code, a, partial);
        if (rc) return rc;
        return fadb_normal_lpdf_finish(shape_code, n, partial, x, mu, sigma, x_adj, mu_adj, sigma_adj,
                                       seed, flags, host_value);
    }
    rc = launch_partial(*c, shape_code, a, partial);
    if (rc) return rc;
    // vector adjoints are done; scalar operands and the value come from the partial sums
    rc = fadb_normal_lpdf_finish(shape_code, n, partial, x, mu, sigma, nullptr, mu_adj, sigma_adj,
                                 seed, flags, host_value);
    if (rc || !guarded || (flags & FADB_STRICT_SIGMA)) return rc;
    // optimistic pass: the undo launch looks at the count the finish kernel published.  A caller that waited for the
    // value already knows whether there is anything to take back (only a -inf / NaN value can come with a count).
    if (host_value && *host_value > -INFINITY) return FADB_OK;
    rc = fadb_normal_lpdf_undo(shape_code, n, x, mu, sigma, x_adj, mu_adj, sigma_adj, seed, flags, invalid);
    if (rc) return rc;
    if (host_value) FADB_CUDA(cudaStreamSynchronize(c->stream));
    return FADB_OK;
}
