// map_reduce.cu — K2/K3: root-level reductions over a VarView with the adjoint scatter fused.
//
// K2 replaces autodiff of ad::sum(f(x)):  SumElemNode::feval/beval (sum.hpp:156-173) over
// UnaryNode::feval/beval (unary.hpp:63-89) or PowNode<n> (pow.hpp:91-162) and the leaf's
// `adj += seed` (var_view.hpp:91-94).  The reference makes ~5 array passes (64-72 B/element);
// here the root seed is known up front so one pass does it: read x (8 B) + RMW x_adj (16 B).
//
// K3 replaces autodiff of ad::prod(x): ProdElemNode::feval/beval (prod.hpp:172-212).  The
// reference recomputes an O(N) product for every exact zero; here pass 1 reduces
// (product of non-zeros, #zeros) and pass 2 scatters adj_i from those two numbers.
#include <type_traits>

#include "common.cuh"
#include "functors.cuh"
#include "p2p_device.cuh"

namespace fadb {

// Link = P2PLink: element-range shard of ONE sum over the ranks of a node.  The adjoints are shard-local
// (the seed is known up front); the last CTA pushes its 1-double partial value into every rank's mailbox over NVLink,
// waits for the other ranks and writes the rank-ordered total: pass + all-reduce in one launch.
template <class F, class Link = NoLink>
__global__ void __launch_bounds__(256)
sum_map_kernel(size_t n, const double* __restrict__ x, double* x_adj, double seed, int overwrite,
               F fun, double* partials, unsigned int* ticket, double* out_value, Link link)
{
    // functors that look values up per thread (the table log) get the table staged in shared memory
    __shared__ double s_tab[F::kTableDoubles ? F::kTableDoubles : 1];
    if (F::kTableDoubles) {
        for (int i = threadIdx.x; i < F::kTableDoubles; i += blockDim.x) s_tab[i] = F::table()[i];
        __syncthreads();
        fun.tab = s_tab;
    }
    pdl_prologue();
    __shared__ double smem[32];
    double acc[1] = {0.0};
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t nth = (size_t)gridDim.x * blockDim.x;
    const bool vec = aligned32(x) && (!x_adj || aligned32(x_adj));
    const size_t n4 = vec ? n / 4 : 0;
    for (size_t q = tid; q < n4; q += nth) {
        const size_t i = 4 * q;
        const double4_t xv = ldg256_stream(x + i);
        double4_t g;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double f = fun.f(xv.v[k]);
            acc[0] += f;
            g.v[k] = fun.b(seed, xv.v[k], f);
        }
        if (x_adj) {
            if (!overwrite) {
                const double4_t old = ld256(x_adj + i);
#pragma unroll
                for (int k = 0; k < 4; ++k) g.v[k] += old.v[k];
            }
            st256(x_adj + i, g);
        }
    }
    for (size_t i = n4 * 4 + tid; i < n; i += nth) {
        const double f = fun.f(x[i]);
        acc[0] += f;
        if (x_adj) {
            const double g = fun.b(seed, x[i], f);
            x_adj[i] = overwrite ? g : x_adj[i] + g;
        }
    }
    block_sum<1>(acc, smem);
    grid_finish<1>(acc, smem, partials, ticket, [&](double (&tot)[1]) {
        if constexpr (std::is_same<Link, P2PLink>::value) {
            p2p_push1(link, tot, 1);
            p2p_pull1(link, tot, 1);
        }
        out_value[0] = tot[0];
    });
}

template <int OP>
struct UnaryFun {
    static constexpr int kTableDoubles = 0;
    const double* tab;
    __device__ static const double* table() { return nullptr; }
    __device__ __forceinline__ double f(double x) const { return Unary<OP>::f(x); }
    __device__ __forceinline__ double b(double s, double x, double f) const { return Unary<OP>::b(s, x, f); }
};
// log through the 129-entry table of fastmath.cuh (15 instead of 27 FP64 instructions, no division, <= 1 ulp)
// for normal positive arguments, libdevice otherwise; adjoint seed / x (unary.hpp:250-255)
struct LogTFun {
    static constexpr int kTableDoubles = 387;
    const double* tab;
    __device__ static const double* table() { return fm::kLogT_d; }
    __device__ __forceinline__ double f(double x) const
    {
        const unsigned long long ux = (unsigned long long)__double_as_longlong(x);
        return (ux - 0x0010000000000000ull < 0x7fe0000000000000ull) ? fm::log_t(x, tab) : log(x);
    }
    __device__ __forceinline__ double b(double s, double x, double) const { return div_guarded(s, x); }
};
struct PowFun {
    static constexpr int kTableDoubles = 0;
    const double* tab;
    __device__ static const double* table() { return nullptr; }
    long n;
    __device__ __forceinline__ double f(double x) const { return pow_int(n, x); }
    __device__ __forceinline__ double b(double s, double x, double f) const { return pow_bmap(n, s, x, f); }
};
struct Pow2Fun {   // pow<2>: f = x*x, adj = 2*seed*(x==0 ? 0 : f/x)   (pow.hpp:125-140)
    static constexpr int kTableDoubles = 0;
    const double* tab;
    __device__ static const double* table() { return nullptr; }
    __device__ __forceinline__ double f(double x) const { return x * x; }
    __device__ __forceinline__ double b(double s, double x, double f) const { return 2. * s * (x == 0 ? 0 : div_guarded(f, x)); }
};

template <class F>
static int launch_sum_map(Context& c, size_t n, const double* x, double* x_adj, double seed,
                          unsigned flags, F fun, double* dval, bool exchange = false)
{
    const int threads = 256;
    if (exchange) {
        P2PLink link;
        int rc = p2p_next_link(c, &link);
        if (rc) return rc;
        const int grid = resident_grid(c, sum_map_kernel<F, P2PLink>, threads, 0, (n + 3) / 4, threads);
        FADB_CUDA(launch_pdl(sum_map_kernel<F, P2PLink>, grid, threads, 0, c.stream, n, x, x_adj, seed,
                             (flags & FADB_OVERWRITE_ADJ) ? 1 : 0, fun, c.partials, c.tickets + 2, dval, link));
    } else {
        const int grid = resident_grid(c, sum_map_kernel<F, NoLink>, threads, 0, (n + 3) / 4, threads);
        FADB_CUDA(launch_pdl(sum_map_kernel<F, NoLink>, grid, threads, 0, c.stream, n, x, x_adj, seed,
                             (flags & FADB_OVERWRITE_ADJ) ? 1 : 0, fun, c.partials, c.tickets + 2, dval, NoLink{}));
    }
    c.launches++;
    FADB_CUDA(cudaGetLastError());
    return FADB_OK;
}

// ------------------------------------------------------------------ prod
__global__ void __launch_bounds__(256)
prod_reduce_kernel(size_t n, const double* __restrict__ x, double* partials, unsigned int* ticket,
                   double* out /* {prod of non-zeros, #zeros, value} */)
{
    pdl_prologue();
    __shared__ double sp[32], sz[32];
    __shared__ bool is_last;
    double p = 1.0, z = 0.0;
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t nth = (size_t)gridDim.x * blockDim.x;
    // read-only stream: 256-bit loads, four in flight per thread (two independent product chains)
    const size_t n4 = aligned32(x) ? n / 4 : 0;
    double p2 = 1.0;
    size_t q = tid;
    for (; q + 3 * nth < n4; q += 4 * nth) {
        double4_t v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ldg256_stream(x + 4 * (q + u * nth));
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double w = v[u].v[k];
                if (w == 0) z += 1.0;
                else if (u & 1) p2 *= w;
                else p *= w;
            }
    }
    for (; q < n4; q += nth) {
        const double4_t v = ldg256_stream(x + 4 * q);
#pragma unroll
        for (int k = 0; k < 4; ++k) { if (v.v[k] == 0) z += 1.0; else p *= v.v[k]; }
    }
    p *= p2;
    for (size_t i = n4 * 4 + tid; i < n; i += nth) {
        const double v = __ldg(x + i);
        if (v == 0) z += 1.0; else p *= v;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    p = warp_prod(p); z = warp_sum(z);
    if (lane == 0) { sp[warp] = p; sz[warp] = z; }
    __syncthreads();
    if (warp == 0) {
        p = warp_prod(lane < nwarp ? sp[lane] : 1.0);
        z = warp_sum(lane < nwarp ? sz[lane] : 0.0);
    }
    if (threadIdx.x == 0) {
        partials[2 * blockIdx.x] = p;
        partials[2 * blockIdx.x + 1] = z;
        __threadfence();
        is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    p = 1.0; z = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
        p *= __ldcg(&partials[2 * b]);
        z += __ldcg(&partials[2 * b + 1]);
    }
    p = warp_prod(p); z = warp_sum(z);
    if (lane == 0) { sp[warp] = p; sz[warp] = z; }
    __syncthreads();
    if (warp == 0) {
        p = warp_prod(lane < nwarp ? sp[lane] : 1.0);
        z = warp_sum(lane < nwarp ? sz[lane] : 0.0);
        if (lane == 0) {
            out[0] = p; out[1] = z; out[2] = (z > 0) ? 0.0 : p;   // prod.hpp:178
            *ticket = 0u;
        }
    }
}

__global__ void __launch_bounds__(256)
prod_scatter_kernel(size_t n, const double* __restrict__ x, double* x_adj, double seed, int overwrite,
                    const double* red)
{
    pdl_prologue();
    const double pnz = red[0], zeros = red[1], val = red[2];
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t nth = (size_t)gridDim.x * blockDim.x;
    // prod.hpp:194-207: zero element -> product of all the others; else seed * (prod / x_i)
    auto adj = [&](double v) { return (v == 0) ? seed * ((zeros > 1) ? 0.0 : pnz) : seed * div_guarded(val, v); };
    const size_t n4 = (aligned32(x) && aligned32(x_adj)) ? n / 4 : 0;
    // BACKWARDS over x: the reduce pass that ran just before swept x upwards, so the top of x is what the L2 still holds
    // (126 MB of it at 2^24 elements = 128 MiB): this pass starts where that one ended
    for (size_t qf = tid; qf < n4; qf += nth) {
        const size_t q = n4 - 1 - qf;
        const double4_t v = ldg256_stream(x + 4 * q);
        double4_t g;
#pragma unroll
        for (int k = 0; k < 4; ++k) g.v[k] = adj(v.v[k]);
        if (!overwrite) {
            const double4_t old = ld256(x_adj + 4 * q);
#pragma unroll
            for (int k = 0; k < 4; ++k) g.v[k] += old.v[k];
        }
        st256(x_adj + 4 * q, g);
    }
    for (size_t i = n4 * 4 + tid; i < n; i += nth) {
        const double g = adj(__ldg(x + i));
        x_adj[i] = overwrite ? g : x_adj[i] + g;
    }
}

}  // namespace fadb

using namespace fadb;

static int read_back(Context& c, const double* dval, double* host_value)
{
    if (!host_value) return FADB_OK;
    FADB_CUDA(cudaMemcpyAsync(c.host_scalars + 1, dval, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    FADB_CUDA(cudaStreamSynchronize(c.stream));
    *host_value = c.host_scalars[1];
    return FADB_OK;
}

static int sum_unary_launch(int op, size_t n, const double* x, double* x_adj, double seed, unsigned flags, double* dev_value,
                            bool exchange)
{
    FADB_REQUIRE(n == 0 || x, "fadb_sum_unary: null x");
    FADB_REQUIRE(dev_value != nullptr, "fadb_sum_unary: null value slot");
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    if (op >= 1000 - 64 && op <= 1000 + 64) {
        const long pw = op - 1000;
        if (pw == 2) return launch_sum_map(*c, n, x, x_adj, seed, flags, Pow2Fun{}, dev_value, exchange);
        return launch_sum_map(*c, n, x, x_adj, seed, flags, PowFun{nullptr, pw}, dev_value, exchange);
    }
    if (op == FADB_LOG) return launch_sum_map(*c, n, x, x_adj, seed, flags, LogTFun{}, dev_value, exchange);
    switch (op) {
#define C(OP) case OP: return launch_sum_map(*c, n, x, x_adj, seed, flags, UnaryFun<OP>{}, dev_value, exchange);
        C(FADB_NEG) C(FADB_SIN) C(FADB_COS) C(FADB_TAN) C(FADB_ASIN) C(FADB_ACOS) C(FADB_ATAN)
        C(FADB_EXP) C(FADB_LOG) C(FADB_SQRT) C(FADB_ERF) C(FADB_SIGMOID) C(FADB_SINH) C(FADB_COSH)
        C(FADB_TANH) C(FADB_MOCK2X) C(FADB_SQUARE)
#undef C
    }
    set_error("fadb_sum_unary: unknown op code " + std::to_string(op));
    return FADB_ERR_ARG;
}

extern "C" int fadb_sum_unary_async(int op, size_t n, const double* x, double* x_adj, double seed,
                                    unsigned flags, double* dev_value)
{
    return sum_unary_launch(op, n, x, x_adj, seed, flags, dev_value, false);
}

// The shard's pass with the all-reduce of the value fused in (csrc/p2p.cu must be connected): dev_value receives
// the sum over ALL ranks' elements, identical bits on every rank; one launch, no collective call.
extern "C" int fadb_sum_unary_allreduce_async(int op, size_t n, const double* x, double* x_adj, double seed,
                                              unsigned flags, double* dev_value)
{
    return sum_unary_launch(op, n, x, x_adj, seed, flags, dev_value, true);
}

extern "C" int fadb_sum_unary(int op, size_t n, const double* x, double* x_adj, double seed,
                              unsigned flags, double* host_value)
{
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    double* dval = c->scalars + 1;
    rc = fadb_sum_unary_async(op, n, x, x_adj, seed, flags, dval);
    if (rc) return rc;
    return read_back(*c, dval, host_value);
}

extern "C" int fadb_prod(size_t n, const double* x, double* x_adj, double seed, unsigned flags,
                         double* host_value)
{
    FADB_REQUIRE(n == 0 || x, "fadb_prod: null x");
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    double* red = c->scalars + 16;   // 3 doubles
    const int threads = 256;
    const int grid = resident_grid(*c, prod_reduce_kernel, threads, 0, n, threads);
    FADB_CUDA(launch_pdl(prod_reduce_kernel, grid, threads, 0, c->stream, n, x, c->partials, c->tickets + 3, red));
    c->launches++;
    FADB_CUDA(cudaGetLastError());
    if (x_adj && n > 0) {
        FADB_CUDA(launch_pdl(prod_scatter_kernel, grid, threads, 0, c->stream, n, x, x_adj, seed, (flags & FADB_OVERWRITE_ADJ) ? 1 : 0,
                             static_cast<const double*>(red)));
        c->launches++;
        FADB_CUDA(cudaGetLastError());
    }
    return read_back(*c, red + 2, host_value);
}
