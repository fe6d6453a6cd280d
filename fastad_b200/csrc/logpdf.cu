// logpdf.cu — fused value + adjoint kernels for the other diagonal adjusted log-pdfs with leaf
// operands: CauchyAdjLogPDFNode (reverse/stat/cauchy.hpp:102-451),
// UniformAdjLogPDFNode (uniform.hpp:101-536), BernoulliAdjLogPDFNode (bernoulli.hpp:94-280).
// Same skeleton as normal.cu: the reference makes one pass per quantity; here one streaming pass
// produces the value sums and every adjoint (256-bit accesses, deterministic reduction).  When
// validity is a property of ALL elements (uniform: min < x < max everywhere; bernoulli; a vector
// cauchy scale), "out of range => no adjoint update" (e.g. cauchy.hpp:151, uniform.hpp:159,
// bernoulli.hpp:145) is reached as in normal.cu: the pass is optimistic (an out-of-range element adds
// nothing and is counted), and lp_undo_kernel behind it takes the vector adjoints back when the count is
// non-zero -- it returns on its first load otherwise.  FADB_STRICT_SIGMA decides validity in a read-only
// pre-pass instead (bit-exact "untouched"), FADB_TRUST_SIGMA drops the guard.
#include "common.cuh"
#include "functors.cuh"

namespace fadb {

struct LpArgs {
    size_t n;
    const double* x; const double* b; const double* c;     // c unused for bernoulli
    double* x_adj; double* b_adj; double* c_adj;
    double seed;
    int overwrite;
    const double* invalid_in;
};

// per-element forward term, validity and adjoints; formulas in the order the reference writes them
template <int DIST>
struct Lp;

template <>
struct Lp<FADB_CAUCHY> {       // x, loc, scale
    static constexpr bool has_c = true, x_has_adj = true;
    __device__ __forceinline__ static bool valid(double, double, double g) { return g > 0; }
    __device__ __forceinline__ static double term(double x, double x0, double g, bool single, const double* = nullptr)
    {
        const double d = x - x0;
        return single ? -log_guarded(g + (d * d) / g) : -log_guarded(g + (1. / g) * (d * d));   // cauchy.hpp:144-146, 216-217
    }
    __device__ __forceinline__ static void adj(double x, double x0, double g, double sd, bool single, double& gx, double& gb, double& gc)
    {
        const double d = x - x0;
        if (single) {                                                   // cauchy.hpp:149-165
            const double inner = g + (d * d) / g;
            const double x0_adj = 2. * d / (g * inner);
            gc = sd * (1. / g * (x0_adj * d - 1));
            gb = sd * x0_adj;
            gx = sd * -x0_adj;
        } else {                                                        // cauchy.hpp:220-237 as written (seed twice)
            const double dx = (-2. * sd) * d / (g * g + d * d);
            gx = dx;
            gb = (-sd) * dx;
            gc = (-sd / g) * (dx * d + 1);
        }
    }
};
template <>
struct Lp<FADB_UNIFORM> {      // x, min, max
    static constexpr bool has_c = true, x_has_adj = false;
    __device__ __forceinline__ static bool valid(double x, double lo, double hi) { return lo < x && x < hi; }
    __device__ __forceinline__ static double term(double, double lo, double hi, bool, const double* = nullptr) { return -log_guarded(hi - lo); }
    __device__ __forceinline__ static void adj(double, double lo, double hi, double sd, bool, double& gx, double& gb, double& gc)
    {
        const double a = sd / (hi - lo);                                // uniform.hpp:157-163
        gx = 0.0; gb = a; gc = -a;
    }
};
template <>
struct Lp<FADB_BERNOULLI> {    // x, p
    static constexpr bool has_c = false, x_has_adj = false;
    __device__ __forceinline__ static bool valid(double x, double p, double) { return 0 < p && p < 1 && (x == 0 || x == 1); }
    // lt: the 129-entry log table of fastmath.cuh in shared memory (15 instead of 27 FP64 instructions, no division, <= 1 ulp):
    // with 32 B/element this kernel is the one log-pdf whose FP64 work exceeds its memory time
    __device__ __forceinline__ static double term(double x, double p, double, bool, const double* lt = nullptr)
    {                                                                   // bernoulli.hpp:119-141
        if (!(0 < p && p < 1)) return (p <= 0) ? (x == 0 ? 0.0 : -INFINITY) : (x == 1 ? 0.0 : -INFINITY);
        if (x != 0 && x != 1) return -INFINITY;
        const double a = x == 0 ? 1 - p : p;               // one log per element, branch-free select; a in (0, 1): normal, positive
        return (lt && a >= 2.3e-308) ? fm::log_t(a, lt) : log_guarded(a);
    }
    __device__ __forceinline__ static void adj(double x, double p, double, double sd, bool, double& gx, double& gb, double& gc)
    {
        gx = 0.0; gc = 0.0;
        gb = div_guarded(x == 0 ? -sd : sd, x == 0 ? 1 - p : p);       // bernoulli.hpp:143-149, one division (Newton reciprocal, <= 1 ulp)
    }
};

// read-only validity pre-pass: #elements that are out of range
template <int DIST, bool B_VEC, bool C_VEC>
__global__ void __launch_bounds__(512)
lp_check_kernel(LpArgs a, double* partials, unsigned int* ticket, double* out_count)
{
    pdl_prologue();
    __shared__ double smem[32];
    double acc[1] = {0.0};
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t nth = (size_t)gridDim.x * blockDim.x;
    const double bs = B_VEC ? 0.0 : __ldg(a.b), cs = (C_VEC || !Lp<DIST>::has_c) ? 0.0 : __ldg(a.c);
    for (size_t i = tid; i < a.n; i += nth) {
        const double x = __ldg(a.x + i), b = B_VEC ? __ldg(a.b + i) : bs, c = C_VEC ? __ldg(a.c + i) : cs;
        acc[0] += Lp<DIST>::valid(x, b, c) ? 0.0 : 1.0;
    }
    block_sum<1>(acc, smem);
    grid_finish<1>(acc, smem, partials, ticket, [&](double (&tot)[1]) { out_count[0] = tot[0]; });
}

__device__ __forceinline__ void lp_rmw4(double* p, const double (&g)[4], int overwrite)
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

// channels: 0 value sum, 1 scalar-b adjoint sum, 2 scalar-c adjoint sum, 3 #invalid
// VAR (bit 0: the old vector adjoints are loaded at the top of the trip, next to the operands, instead of at the
// read-modify-write; bit 1: the NEXT trip's operands are loaded before this trip's arithmetic, register double-buffered)
template <int DIST, bool B_VEC, bool C_VEC, int VAR>
__global__ void __launch_bounds__(256, (VAR & 2) ? 3 : 4)
lp_kernel(LpArgs a, double* partials, unsigned int* ticket, double* out /* [4] value, invalid, -, - */)
{
    constexpr bool EARLY = VAR & 1, PIPE = VAR & 2;
    __shared__ double s_logt[DIST == FADB_BERNOULLI ? 387 : 1];
    if (DIST == FADB_BERNOULLI) {           // constant data: copied before the dependency wait
        for (int i = threadIdx.x; i < 387; i += blockDim.x) s_logt[i] = fm::kLogT_d[i];
        __syncthreads();
    }
    const double* lt = DIST == FADB_BERNOULLI ? s_logt : nullptr;
    pdl_prologue();
    __shared__ double smem[32 * 4];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t nth = (size_t)gridDim.x * blockDim.x;
    const bool single = a.n == 1;
    const double bs = B_VEC ? 0.0 : __ldg(a.b), cs = (C_VEC || !Lp<DIST>::has_c) ? 0.0 : __ldg(a.c);
    bool valid_all = a.invalid_in ? (__ldcg(a.invalid_in) == 0.0) : true;
    if (DIST == FADB_CAUCHY && !C_VEC) valid_all = cs > 0;    // a scalar scale: every thread sees it (cauchy.hpp:168-170)
    const bool back = valid_all && a.seed != 0.0;
    const bool wx = back && Lp<DIST>::x_has_adj && a.x_adj, wb = back && a.b_adj, wc = back && Lp<DIST>::has_c && a.c_adj;

    const bool vec = aligned32(a.x) && (!B_VEC || aligned32(a.b)) && (!C_VEC || aligned32(a.c)) &&
                     (!a.x_adj || aligned32(a.x_adj)) && (!B_VEC || !a.b_adj || aligned32(a.b_adj)) &&
                     (!C_VEC || !a.c_adj || aligned32(a.c_adj));
    const size_t n4 = vec ? a.n / 4 : 0;
    if (VAR == 0) {
    for (size_t q = tid; q < n4; q += nth) {
        const size_t i = 4 * q;
        const double4_t xv = ldg256_stream(a.x + i);
        double4_t bv, cv;
        if (B_VEC) bv = ldg256_stream(a.b + i);
        if (C_VEC) cv = ldg256_stream(a.c + i);
        double gx[4], gb[4], gc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double b = B_VEC ? bv.v[k] : bs, c = C_VEC ? cv.v[k] : cs;
            const bool ok = Lp<DIST>::valid(xv.v[k], b, c);
            if (!ok) acc[3] += 1.0;
            acc[0] += (ok || DIST == FADB_BERNOULLI) ? Lp<DIST>::term(xv.v[k], b, c, single, lt) : 0.0;
            Lp<DIST>::adj(xv.v[k], b, c, a.seed, single, gx[k], gb[k], gc[k]);
            if (!ok) gx[k] = gb[k] = gc[k] = 0.0;          // optimistic pass: an out-of-range element adds nothing
            if (!B_VEC) acc[1] += gb[k];
            if (!C_VEC) acc[2] += gc[k];
        }
        if (wx) lp_rmw4(a.x_adj + i, gx, a.overwrite);
        if (wb && B_VEC) lp_rmw4(a.b_adj + i, gb, a.overwrite);
        if (wc && C_VEC) lp_rmw4(a.c_adj + i, gc, a.overwrite);
    }
    } else {
    double4_t xv, bv, cv;
    if (PIPE && tid < n4) {
        xv = ldg256_stream(a.x + 4 * tid);
        if (B_VEC) bv = ldg256_stream(a.b + 4 * tid);
        if (C_VEC) cv = ldg256_stream(a.c + 4 * tid);
    }
    for (size_t q = tid; q < n4; q += nth) {
        const size_t i = 4 * q;
        double4_t xn, bn, cn;
        if (PIPE) {
            if (q + nth < n4) {
                xn = ldg256_stream(a.x + i + 4 * nth);
                if (B_VEC) bn = ldg256_stream(a.b + i + 4 * nth);
                if (C_VEC) cn = ldg256_stream(a.c + i + 4 * nth);
            }
        } else {
            xv = ldg256_stream(a.x + i);
            if (B_VEC) bv = ldg256_stream(a.b + i);
            if (C_VEC) cv = ldg256_stream(a.c + i);
        }
        double4_t ox, ob, oc;           // old adjoints, in flight during the arithmetic
        if (EARLY && !a.overwrite) {
            if (wx) ox = ld256(a.x_adj + i);
            if (wb && B_VEC) ob = ld256(a.b_adj + i);
            if (wc && C_VEC) oc = ld256(a.c_adj + i);
        }
        double gx[4], gb[4], gc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double b = B_VEC ? bv.v[k] : bs, c = C_VEC ? cv.v[k] : cs;
            const bool ok = Lp<DIST>::valid(xv.v[k], b, c);
            if (!ok) acc[3] += 1.0;
            acc[0] += (ok || DIST == FADB_BERNOULLI) ? Lp<DIST>::term(xv.v[k], b, c, single, lt) : 0.0;
            Lp<DIST>::adj(xv.v[k], b, c, a.seed, single, gx[k], gb[k], gc[k]);
            if (!ok) gx[k] = gb[k] = gc[k] = 0.0;          // optimistic pass: an out-of-range element adds nothing
            if (!B_VEC) acc[1] += gb[k];
            if (!C_VEC) acc[2] += gc[k];
        }
        if (EARLY && !a.overwrite) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { ox.v[k] += gx[k]; ob.v[k] += gb[k]; oc.v[k] += gc[k]; }
            if (wx) st256(a.x_adj + i, ox);
            if (wb && B_VEC) st256(a.b_adj + i, ob);
            if (wc && C_VEC) st256(a.c_adj + i, oc);
        } else {
            if (wx) lp_rmw4(a.x_adj + i, gx, a.overwrite);
            if (wb && B_VEC) lp_rmw4(a.b_adj + i, gb, a.overwrite);
            if (wc && C_VEC) lp_rmw4(a.c_adj + i, gc, a.overwrite);
        }
        if (PIPE) { xv = xn; bv = bn; cv = cn; }
    }
    }
    for (size_t i = n4 * 4 + tid; i < a.n; i += nth) {
        const double x = a.x[i], b = B_VEC ? a.b[i] : bs, c = C_VEC ? a.c[i] : cs;
        const bool ok = Lp<DIST>::valid(x, b, c);
        if (!ok) acc[3] += 1.0;
        acc[0] += (ok || DIST == FADB_BERNOULLI) ? Lp<DIST>::term(x, b, c, single, lt) : 0.0;
        double gx, gb, gc;
        Lp<DIST>::adj(x, b, c, a.seed, single, gx, gb, gc);
        if (!ok) gx = gb = gc = 0.0;
        if (!B_VEC) acc[1] += gb;
        if (!C_VEC) acc[2] += gc;
        if (wx) a.x_adj[i] = a.overwrite ? gx : a.x_adj[i] + gx;
        if (wb && B_VEC) a.b_adj[i] = a.overwrite ? gb : a.b_adj[i] + gb;
        if (wc && C_VEC) a.c_adj[i] = a.overwrite ? gc : a.c_adj[i] + gc;
    }
    block_sum<4>(acc, smem);
    grid_finish<4>(acc, smem, partials, ticket, [&](double (&tot)[4]) {
        const bool invalid = tot[3] != 0.0 || !valid_all;
        out[1] = invalid ? 1.0 : 0.0;
        out[0] = (DIST == FADB_BERNOULLI) ? tot[0] : (invalid ? -INFINITY : tot[0]);
        // adjoints of scalar operands (summed over the elements), only when everything is in range
        if (!invalid && a.seed != 0.0) {
            if (!B_VEC && a.b_adj) a.b_adj[0] = a.overwrite ? tot[1] : a.b_adj[0] + tot[1];
            if (!C_VEC && Lp<DIST>::has_c && a.c_adj) a.c_adj[0] = a.overwrite ? tot[2] : a.c_adj[0] + tot[2];
        }
    });
}

// undo of the optimistic pass: `verdict` = out[1] of lp_kernel (1 = something was out of range).  Zero: every thread
// leaves after one load.  Otherwise the contributions are recomputed as lp_kernel formed them and taken back.
template <int DIST, bool B_VEC, bool C_VEC>
__global__ void __launch_bounds__(256)
lp_undo_kernel(LpArgs a, const double* verdict)
{
    pdl_prologue();
    if (__ldcg(verdict) == 0.0 || a.seed == 0.0) return;
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t nth = (size_t)gridDim.x * blockDim.x;
    const bool single = a.n == 1;
    const double bs = B_VEC ? 0.0 : __ldg(a.b), cs = (C_VEC || !Lp<DIST>::has_c) ? 0.0 : __ldg(a.c);
    const bool wx = Lp<DIST>::x_has_adj && a.x_adj, wb = B_VEC && a.b_adj, wc = C_VEC && Lp<DIST>::has_c && a.c_adj;
    for (size_t i = tid; i < a.n; i += nth) {
        const double x = a.x[i], b = B_VEC ? a.b[i] : bs, c = C_VEC ? a.c[i] : cs;
        double gx, gb, gc;
        Lp<DIST>::adj(x, b, c, a.seed, single, gx, gb, gc);
        if (!Lp<DIST>::valid(x, b, c)) gx = gb = gc = 0.0;
        if (wx) a.x_adj[i] = a.overwrite ? 0.0 : a.x_adj[i] - gx;
        if (wb) a.b_adj[i] = a.overwrite ? 0.0 : a.b_adj[i] - gb;
        if (wc) a.c_adj[i] = a.overwrite ? 0.0 : a.c_adj[i] - gc;
    }
}

template <int DIST, bool B_VEC, bool C_VEC>
static int run_logpdf(Context& c, LpArgs a, unsigned flags, bool any_adj, double* out)
{
    // validity visible to every thread only for a cauchy with a scalar scale
    const bool global_check = !(DIST == FADB_CAUCHY && !C_VEC);
    double* invalid = out + 2;
    a.invalid_in = nullptr;
    // only VECTOR adjoints are written by the pass itself (scalar ones by its epilogue, which knows the verdict)
    const bool vec_adj = (Lp<DIST>::x_has_adj && a.x_adj) || (B_VEC && a.b_adj) || (C_VEC && Lp<DIST>::has_c && a.c_adj);
    const bool guarded = any_adj && vec_adj && a.seed != 0.0 && global_check && !(flags & FADB_TRUST_SIGMA);
    if (guarded && (flags & FADB_STRICT_SIGMA)) {
        const int threads = 512;
        const int grid = resident_grid(c, lp_check_kernel<DIST, B_VEC, C_VEC>, threads, 0, a.n, threads);
        FADB_CUDA(launch_pdl(lp_check_kernel<DIST, B_VEC, C_VEC>, grid, threads, 0, c.stream, a, c.partials, c.tickets + 9, invalid));
        c.launches++;
        a.invalid_in = invalid;
    }
    const int threads = 256;
    // Bernoulli (32 B per element, the one log-pdf whose arithmetic is comparable to its memory time): ncu showed two exposed
    // waits per trip, on the operands and on the old adjoints of the read-modify-write (25 % + 22 % of all stall samples).  Loading
    // the old adjoints next to the operands leaves one: 0.206 -> 0.199 ms at 2^25 elements; double-buffering the operands as well
    // (VAR 3, 80 registers) measured the same.  FADB_LP_VARIANT=0|3|5 selects a variant for A/B runs.
    static const int var_env = getenv("FADB_LP_VARIANT") ? atoi(getenv("FADB_LP_VARIANT")) : -1;
    const int var = DIST == FADB_BERNOULLI ? (var_env >= 0 ? (var_env & 7) : 5) : 0;
#define FADB_LP_LAUNCH(V)                                                                                                       \
    do {                                                                                                                        \
        const int grid = resident_grid(c, lp_kernel<DIST, B_VEC, C_VEC, V>, threads, 0, (a.n + 3) / 4, threads);                \
        FADB_CUDA(launch_pdl(lp_kernel<DIST, B_VEC, C_VEC, V>, grid, threads, 0, c.stream, a, c.partials, c.tickets + 10, out)); \
    } while (0)
    if (DIST == FADB_BERNOULLI && var == 5) FADB_LP_LAUNCH(DIST == FADB_BERNOULLI ? 5 : 0);
    else if (DIST == FADB_BERNOULLI && var == 3) FADB_LP_LAUNCH(DIST == FADB_BERNOULLI ? 3 : 0);
    else FADB_LP_LAUNCH(0);
#undef FADB_LP_LAUNCH
    c.launches++;
    if (guarded && !(flags & FADB_STRICT_SIGMA)) {
        const int ugrid = resident_grid(c, lp_undo_kernel<DIST, B_VEC, C_VEC>, threads, 0, a.n, threads);
        FADB_CUDA(launch_pdl(lp_undo_kernel<DIST, B_VEC, C_VEC>, ugrid, threads, 0, c.stream, a, out + 1));
        c.launches++;
    }
    FADB_CUDA(cudaGetLastError());
    return FADB_OK;
}

}  // namespace fadb

using namespace fadb;

// value left in dev_out[0] (device, 4 doubles of scratch); asynchronous
extern "C" int fadb_logpdf_async(int dist, int shape_code, size_t n, const double* x, const double* b,
                                 const double* c, double* x_adj, double* b_adj, double* c_adj, double seed,
                                 unsigned flags, double* dev_out4)
{
    FADB_REQUIRE(dist >= FADB_CAUCHY && dist <= FADB_BERNOULLI, "fadb_logpdf: unknown distribution");
    FADB_REQUIRE(shape_code >= 0 && shape_code <= 3 && dev_out4, "fadb_logpdf: bad shape_code / null output");
    FADB_REQUIRE(n == 0 || (x && b && (dist == FADB_BERNOULLI || c)), "fadb_logpdf: null operand");
    Context* ctx;
    int rc = current_context(&ctx);
    if (rc) return rc;
    LpArgs a{n, x, b, c, x_adj, b_adj, c_adj, seed, (flags & FADB_OVERWRITE_ADJ) ? 1 : 0, nullptr};
    const bool any_adj = x_adj || b_adj || c_adj;
    const bool bv = shape_code & 1, cv = shape_code & 2;
#define FADB_LP(D)                                                                              \
    do {                                                                                        \
        if (bv && cv) return run_logpdf<D, true, true>(*ctx, a, flags, any_adj, dev_out4);     \
        if (bv) return run_logpdf<D, true, false>(*ctx, a, flags, any_adj, dev_out4);          \
        if (cv) return run_logpdf<D, false, true>(*ctx, a, flags, any_adj, dev_out4);          \
        return run_logpdf<D, false, false>(*ctx, a, flags, any_adj, dev_out4);                 \
    } while (0)
    if (dist == FADB_CAUCHY) FADB_LP(FADB_CAUCHY);
    if (dist == FADB_UNIFORM) FADB_LP(FADB_UNIFORM);
    if (bv) return run_logpdf<FADB_BERNOULLI, true, false>(*ctx, a, flags, any_adj, dev_out4);
    return run_logpdf<FADB_BERNOULLI, false, false>(*ctx, a, flags, any_adj, dev_out4);
#undef FADB_LP
}

extern "C" int fadb_logpdf(int dist, int shape_code, size_t n, const double* x, const double* b,
                           const double* c, double* x_adj, double* b_adj, double* c_adj, double seed,
                           unsigned flags, double* host_value)
{
    Context* ctx;
    int rc = current_context(&ctx);
    if (rc) return rc;
    double* out = ctx->scalars + 24;
    rc = fadb_logpdf_async(dist, shape_code, n, x, b, c, x_adj, b_adj, c_adj, seed, flags, out);
    if (rc) return rc;
    if (host_value) {
        FADB_CUDA(cudaMemcpyAsync(ctx->host_scalars + 4, out, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        FADB_CUDA(cudaStreamSynchronize(ctx->stream));
        *host_value = ctx->host_scalars[4];
    }
    return FADB_OK;
}
