// fastad_b200/fused.cuh — type-fused expression kernels (nvcc only).
//
// FastAD gets loop fusion from its C++ expression TYPE (every node's feval/beval is inlined
// into one expression, unary.hpp:63-89, binary.hpp:100-136).  A pre-built shared library cannot
// instantiate user types, so the C ABI falls back to an interpreter for shapes it has no
// hand-written kernel for (expr.cu; measured ~8 % of the HBM roofline: node values live in local
// memory).  When the user's translation unit is compiled with nvcc, this header does what the
// reference does: the expression TYPE instantiates the kernel.  Each node type gets a device
// view `Dev<Node>` whose fwd()/bwd() are the node's feval()/beval() for ONE element, node values
// live in a per-thread `Tape` struct (registers), and a hand-written grid-stride skeleton
// (`fused_kernel`) runs forward + adjoint sweep per element with the same deterministic
// reduction epilogue as the library kernels.
//
//   auto f = ad::fuse(expr);            // instead of ad::bind(expr)
//   double v = ad::autodiff(f);         // scalar root (sum / normal_adj_log_pdf / scalar expr)
//   ad::autodiff(f, ad::device_seed{p}) // vector root, seed and value stay on the device
//
// Fusable: VarView/Var, constants, unary, binary, pow<n>, placeholders + glue, with an optional
// ROOT reduction ad::sum(expr) or ad::normal_adj_log_pdf(x, mu, sigma) over elementwise operands
// (vector sigma as a variable needs ad::fuse(expr, FADB_TRUST_SIGMA), see normal.hpp:440-457).
// Anything else (dot, prod, sum/prod over expression lists, nested reductions) is rejected at
// compile time: use ad::bind (C ABI).
#pragma once
#ifndef __CUDACC__
#error "fastad_b200/fused.cuh must be compiled with nvcc (-gencode arch=compute_100a,code=sm_100a)"
#endif
#include <cuda_runtime.h>

#include "ad.hpp"
#include "device_math.cuh"

namespace ad {
namespace fused {

constexpr int kMaxCh = 8;          // [0..2] terminal sums, [3..] adjoints of broadcast scalar leaves
struct Acc {
    double ch[kMaxCh];
};
enum : int { T_STORE = 0, T_SUM = 1, T_NORMAL = 2 };

struct Ctx {                       // host-side preparation
    size_t N = 1;
    int next_ch = 3;
    std::vector<double*> chan_adj;                 // adjoint pointer of each channel >= 3
    std::vector<std::shared_ptr<detail::Storage>> storages;
    void touch(const std::shared_ptr<detail::Storage>& s)
    {
        for (auto& t : storages) if (t == s) return;
        storages.push_back(s);
    }
    int channel_for(double* adj)
    {
        for (size_t k = 0; k < chan_adj.size(); ++k) if (chan_adj[k] == adj) return 3 + (int)k;
        if (next_ch >= kMaxCh) throw Error("fastad_b200: fused expression has more than 5 scalar variables in a vector domain");
        chan_adj.push_back(adj);
        return next_ch++;
    }
};

template <class E, class = void>
struct Dev;                        // device view of a node; specialised below
template <class E, class = void>
struct fusable : std::false_type {};

// ------------------------------------------------------------------------------ leaves
template <class S>
struct Dev<VarView<double, S>> {
    double* val;
    double* adj;
    int bcast, chan;
    struct Tape { double v; };
    static Dev make(const VarView<double, S>& e, Ctx& c)
    {
        e.storage()->to_device();
        c.touch(e.storage());
        Dev d{e.data(), e.data_adj(), 0, -1};
        if (e.size() != c.N) {
            if (e.size() != 1) throw Error("fastad_b200: fused expression mixes vector shapes");
            d.bcast = 1;
            if (d.adj) d.chan = c.channel_for(d.adj);
        }
        return d;
    }
    __device__ __forceinline__ void fwd(size_t i, Tape& t, Acc&) const                                // var_view.hpp:69
    {
        t.v = val[bcast ? 0 : i];
        // the adjoint read-modify-write comes after the whole forward pass: start fetching its line
        // now (a prefetch, not a load: the same leaf may occur twice or be a placeholder)
        if (adj && !bcast) asm volatile("prefetch.global.L1 [%0];" ::"l"(adj + i));
    }
    __device__ __forceinline__ void bwd(double s, size_t i, Tape&, Acc& a) const                      // var_view.hpp:91-94
    {
        if (!adj) return;
        if (bcast) a.ch[chan] += s;
        else adj[i] += s;
    }
};
template <class S> struct fusable<VarView<double, S>> : std::true_type {};

template <>
struct Dev<core::Constant<double, scl>> {
    double c;
    struct Tape { double v; };
    static Dev make(const core::Constant<double, scl>& e, Ctx&) { return Dev{e.feval()}; }
    __device__ __forceinline__ void fwd(size_t, Tape& t, Acc&) const { t.v = c; }
    __device__ __forceinline__ void bwd(double, size_t, Tape&, Acc&) const {}                          // constant.hpp:55-56
};
template <> struct fusable<core::Constant<double, scl>> : std::true_type {};

template <class S>
struct Dev<core::ConstantView<double, S>> {
    const double* val;
    int bcast;
    struct Tape { double v; };
    static Dev make(const core::ConstantView<double, S>& e, Ctx& c)
    {
        if (e.size() != c.N && e.size() != 1) throw Error("fastad_b200: fused expression mixes vector shapes");
        return Dev{e.data(), e.size() != c.N};
    }
    __device__ __forceinline__ void fwd(size_t i, Tape& t, Acc&) const { t.v = __ldg(val + (bcast ? 0 : i)); }
    __device__ __forceinline__ void bwd(double, size_t, Tape&, Acc&) const {}
};
template <class S> struct fusable<core::ConstantView<double, S>> : std::true_type {};

template <class S>
struct Dev<core::Constant<double, S>, std::enable_if_t<!util::is_scl_tag_v<S>>> {
    const double* val;
    int bcast;
    struct Tape { double v; };
    static Dev make(const core::Constant<double, S>& e, Ctx& c)
    {
        e.storage()->to_device();
        c.touch(e.storage());
        if (e.size() != c.N && e.size() != 1) throw Error("fastad_b200: fused expression mixes vector shapes");
        return Dev{e.storage()->dval, e.size() != c.N};
    }
    __device__ __forceinline__ void fwd(size_t i, Tape& t, Acc&) const { t.v = __ldg(val + (bcast ? 0 : i)); }
    __device__ __forceinline__ void bwd(double, size_t, Tape&, Acc&) const {}
};
template <class S> struct fusable<core::Constant<double, S>, std::enable_if_t<!util::is_scl_tag_v<S>>> : std::true_type {};

// ------------------------------------------------------------------------------ elementwise nodes
template <class Tag, class E>
struct Dev<core::UnaryNode<Tag, E>> {
    Dev<E> e;
    struct Tape { typename Dev<E>::Tape c; double v; };
    static Dev make(const core::UnaryNode<Tag, E>& n, Ctx& c) { return Dev{Dev<E>::make(n.child(), c)}; }
    __device__ __forceinline__ void fwd(size_t i, Tape& t, Acc& a) const                               // unary.hpp:63-68
    {
        e.fwd(i, t.c, a);
        t.v = fadb::Unary<Tag::code>::f(t.c.v);
    }
    __device__ __forceinline__ void bwd(double s, size_t i, Tape& t, Acc& a) const                     // unary.hpp:81-89
    {
        e.bwd(fadb::Unary<Tag::code>::b(s, t.c.v, t.v), i, t.c, a);
    }
};
template <class Tag, class E> struct fusable<core::UnaryNode<Tag, E>> : fusable<E> {};

template <int Code> struct BinOp;
template <> struct BinOp<FADB_ADD> {
    __device__ __forceinline__ static double f(double x, double y) { return x + y; }
    __device__ __forceinline__ static double bl(double s, double, double, double) { return s; }
    __device__ __forceinline__ static double br(double s, double, double, double) { return s; }
};
template <> struct BinOp<FADB_SUB> {
    __device__ __forceinline__ static double f(double x, double y) { return x - y; }
    __device__ __forceinline__ static double bl(double s, double, double, double) { return s; }
    __device__ __forceinline__ static double br(double s, double, double, double) { return -s; }
};
template <> struct BinOp<FADB_MUL> {
    __device__ __forceinline__ static double f(double x, double y) { return x * y; }
    __device__ __forceinline__ static double bl(double s, double, double y, double) { return s * y; }
    __device__ __forceinline__ static double br(double s, double x, double, double) { return s * x; }
};
template <> struct BinOp<FADB_DIV> {
    __device__ __forceinline__ static double f(double x, double y) { return x / y; }
    __device__ __forceinline__ static double bl(double s, double, double y, double) { return s / y; }
    __device__ __forceinline__ static double br(double s, double, double y, double f) { return -s * f / y; }
};
template <> struct BinOp<FADB_MOCKB> {
    __device__ __forceinline__ static double f(double x, double y) { return x - 2 * y; }
    __device__ __forceinline__ static double bl(double s, double, double, double) { return s; }
    __device__ __forceinline__ static double br(double s, double, double, double) { return -2. * s; }
};

template <class Tag, class L, class R>
struct Dev<core::BinaryNode<Tag, L, R>> {
    Dev<L> l;
    Dev<R> r;
    struct Tape { typename Dev<L>::Tape a; typename Dev<R>::Tape b; double v; };
    static Dev make(const core::BinaryNode<Tag, L, R>& n, Ctx& c)
    {
        return Dev{Dev<L>::make(n.lhs(), c), Dev<R>::make(n.rhs(), c)};
    }
    __device__ __forceinline__ void fwd(size_t i, Tape& t, Acc& a) const                               // binary.hpp:100-106
    {
        l.fwd(i, t.a, a);
        r.fwd(i, t.b, a);
        t.v = BinOp<Tag::code>::f(t.a.v, t.b.v);
    }
    __device__ __forceinline__ void bwd(double s, size_t i, Tape& t, Acc& a) const                     // binary.hpp:120-136
    {
        const double rs = BinOp<Tag::code>::br(s, t.a.v, t.b.v, t.v);
        const double ls = BinOp<Tag::code>::bl(s, t.a.v, t.b.v, t.v);
        r.bwd(rs, i, t.b, a);                                                                          // rhs first
        l.bwd(ls, i, t.a, a);
    }
};
template <class Tag, class L, class R>
struct fusable<core::BinaryNode<Tag, L, R>>
    : std::bool_constant<fusable<L>::value && fusable<R>::value && (Tag::code <= FADB_DIV || Tag::code == FADB_MOCKB)> {};

template <int64_t N, class E>
struct Dev<core::PowNode<N, E>> {
    Dev<E> e;
    struct Tape { typename Dev<E>::Tape c; double v; };
    static Dev make(const core::PowNode<N, E>& n, Ctx& c) { return Dev{Dev<E>::make(n.child(), c)}; }
    __device__ __forceinline__ void fwd(size_t i, Tape& t, Acc& a) const                               // pow.hpp:91-99
    {
        e.fwd(i, t.c, a);
        t.v = fadb::pow_int((long)N, t.c.v);
    }
    __device__ __forceinline__ void bwd(double s, size_t i, Tape& t, Acc& a) const                     // pow.hpp:101-162
    {
        e.bwd(fadb::pow_bmap((long)N, s, t.c.v, t.v), i, t.c, a);
    }
};
template <int64_t N, class E> struct fusable<core::PowNode<N, E>> : fusable<E> {};

// placeholder  w = expr  (eq.hpp:89-107).  Readers of w are ordinary VarView leaves: they load the
// value this thread stored a moment ago and add their seeds to w.adj[i] before this node's bwd.
template <class V, class E>
struct Dev<core::EqNode<V, E>> {
    Dev<V> w;
    Dev<E> e;
    struct Tape { typename Dev<E>::Tape c; double v; };
    static Dev make(const core::EqNode<V, E>& n, Ctx& c)
    {
        if (n.var().size() != c.N) throw Error("fastad_b200: fused placeholders must have the shape of the fused domain");
        return Dev{Dev<V>::make(n.var(), c), Dev<E>::make(n.expr(), c)};
    }
    __device__ __forceinline__ void fwd(size_t i, Tape& t, Acc& a) const
    {
        e.fwd(i, t.c, a);
        t.v = t.c.v;
        w.val[i] = t.v;
    }
    __device__ __forceinline__ void bwd(double s, size_t i, Tape& t, Acc& a) const
    {
        const double tot = w.adj[i] + s;
        w.adj[i] = tot;
        e.bwd(tot, i, t.c, a);
    }
};
template <class V, class E> struct fusable<core::EqNode<V, E>> : fusable<E> {};

template <class L, class R>
struct Dev<core::GlueNode<L, R>> {
    Dev<L> l;
    Dev<R> r;
    struct Tape { typename Dev<L>::Tape a; typename Dev<R>::Tape b; double v; };
    static Dev make(const core::GlueNode<L, R>& n, Ctx& c)
    {
        return Dev{Dev<L>::make(n.lhs(), c), Dev<R>::make(n.rhs(), c)};
    }
    __device__ __forceinline__ void fwd(size_t i, Tape& t, Acc& a) const                               // glue.hpp:67-71
    {
        l.fwd(i, t.a, a);
        r.fwd(i, t.b, a);
        t.v = t.b.v;
    }
    __device__ __forceinline__ void bwd(double s, size_t i, Tape& t, Acc& a) const                     // glue.hpp:81-86
    {
        r.bwd(s, i, t.b, a);
        l.bwd(0.0, i, t.a, a);
    }
};
template <class L, class R>
struct fusable<core::GlueNode<L, R>> : std::bool_constant<fusable<L>::value && fusable<R>::value> {};

// ------------------------------------------------------------------------------ root reductions
template <class E>
struct Dev<core::SumElemNode<E>> {
    Dev<E> e;
    struct Tape { typename Dev<E>::Tape c; double v; };
    static Dev make(const core::SumElemNode<E>& n, Ctx& c) { return Dev{Dev<E>::make(n.child(), c)}; }
    __device__ __forceinline__ void fwd(size_t i, Tape& t, Acc& a) const                               // sum.hpp:156-164
    {
        e.fwd(i, t.c, a);
        t.v = t.c.v;
        a.ch[0] += t.c.v;
    }
    __device__ __forceinline__ void bwd(double s, size_t i, Tape& t, Acc& a) const { e.bwd(s, i, t.c, a); }   // sum.hpp:170-173
};

// NormNode as a root: sum of squares, child seed 2*seed*x (norm.hpp:47-57)
template <class E>
struct Dev<core::NormNode<E>> {
    Dev<E> e;
    struct Tape { typename Dev<E>::Tape c; double v; };
    static Dev make(const core::NormNode<E>& n, Ctx& c) { return Dev{Dev<E>::make(n.child(), c)}; }
    __device__ __forceinline__ void fwd(size_t i, Tape& t, Acc& a) const
    {
        e.fwd(i, t.c, a);
        t.v = t.c.v * t.c.v;
        a.ch[0] += t.v;
    }
    __device__ __forceinline__ void bwd(double s, size_t i, Tape& t, Acc& a) const { e.bwd(s * 2. * t.c.v, i, t.c, a); }
};

template <class X, class M, class S>
struct Dev<core::NormalAdjLogPDFNode<X, M, S>> {
    Dev<X> x;
    Dev<M> m;
    Dev<S> s;
    int mu_vec, sig_vec;
    unsigned long long n;
    struct Tape { typename Dev<X>::Tape tx; typename Dev<M>::Tape tm; typename Dev<S>::Tape ts; double v, d, z; };
    static Dev make(const core::NormalAdjLogPDFNode<X, M, S>& nd, Ctx& c)
    {
        return Dev{Dev<X>::make(nd.x(), c), Dev<M>::make(nd.mean(), c), Dev<S>::make(nd.sigma(), c),
                   nd.mean().size() > 1, nd.sigma().size() > 1, c.N};
    }
    __device__ __forceinline__ void fwd(size_t i, Tape& t, Acc& a) const
    {
        x.fwd(i, t.tx, a);
        m.fwd(i, t.tm, a);
        s.fwd(i, t.ts, a);
        t.d = t.tx.v - t.tm.v;
        const double sg = t.ts.v;
        if (sig_vec) {
            if (!(sg > 0)) a.ch[2] += 1.0;
            t.z = t.d / sg;
            a.ch[0] += t.z * t.z;                                                                      // normal.hpp:449, 546
            a.ch[1] += fadb::log_guarded(sg);
        } else if (n == 1) {
            t.z = t.d / sg;                                                                            // sss, normal.hpp:153-155
            a.ch[0] += t.z * t.z;
        } else {
            a.ch[0] += t.d * t.d;                                                                      // normal.hpp:244
        }
        t.v = 0.0;
    }
    __device__ __forceinline__ void bwd(double sd, size_t i, Tape& t, Acc& a) const
    {
        const double xv = t.tx.v, mv = t.tm.v, sg = t.ts.v, d = t.d;
        double gx, gm, gs;
        if (sig_vec) {
            const double s2 = sg * sg;
            if (mu_vec) { gx = sd * (mv - xv) / s2; gm = sd * d / s2; }                                // normal.hpp:563-564
            else { gx = (sd / s2) * (mv - xv); gm = sd * (d / s2); }                                   // normal.hpp:471-474
            gs = (sd / sg) * (t.z * t.z - 1.);                                                         // normal.hpp:468, 560
        } else if (n == 1) {
            const double inv_s = 1. / sg, z = d * inv_s;                                               // normal.hpp:158-171
            gs = sd * (z * z - 1) * inv_s;
            gm = sd * z * inv_s;
            gx = -gm;
        } else {
            const double inv_s = 1. / sg, inv_s_sq = inv_s * inv_s, c = sd * inv_s_sq;
            gx = c * (mv - xv);                                                                        // normal.hpp:282, 371
            gm = mu_vec ? c * d : sd * (d * inv_s_sq);                                                 // normal.hpp:370, 277
            gs = sd * ((d * d) / (sg * sg) - 1.) * inv_s;                                              // normal.hpp:273
        }
        s.bwd(gs, i, t.ts, a);                                                                         // sigma, mu, x order
        m.bwd(gm, i, t.tm, a);
        x.bwd(gx, i, t.tx, a);
    }
};

// terminal kind of a root: the rightmost statement of a glue chain decides
template <class E> struct terminal { static constexpr int kind = T_STORE; using last_t = E; };
template <class L, class R> struct terminal<core::GlueNode<L, R>> : terminal<R> {};
template <class E> struct terminal<core::SumElemNode<E>> { static constexpr int kind = T_SUM; using last_t = core::SumElemNode<E>; };
template <class E> struct terminal<core::NormNode<E>> { static constexpr int kind = T_SUM; using last_t = core::NormNode<E>; };
template <class X, class M, class S> struct terminal<core::NormalAdjLogPDFNode<X, M, S>> {
    static constexpr int kind = T_NORMAL; using last_t = core::NormalAdjLogPDFNode<X, M, S>;
};
// a root may END in a reduction; reductions anywhere else are not fusable
template <class E> struct root_fusable : fusable<E> {};
template <class E> struct root_fusable<core::SumElemNode<E>> : fusable<E> {};
template <class E> struct root_fusable<core::NormNode<E>> : fusable<E> {};
template <class X, class M, class S>
struct root_fusable<core::NormalAdjLogPDFNode<X, M, S>>
    : std::bool_constant<fusable<X>::value && fusable<M>::value && fusable<S>::value> {};
template <class L, class R>
struct root_fusable<core::GlueNode<L, R>> : std::bool_constant<fusable<L>::value && root_fusable<R>::value> {};

template <class E> __device__ __forceinline__ const E& last_of(const E& d) { return d; }
template <class L, class R>
__device__ __forceinline__ const auto& last_of(const Dev<core::GlueNode<L, R>>& d) { return last_of(d.r); }

// ------------------------------------------------------------------------------ kernel skeleton
struct Args {
    unsigned long long N;
    int mode;                  // 1 forward, 2 backward, 3 both
    double seed;
    const double* seed_vec;
    double* out_vec;           // T_STORE: per-element value (may be null)
    double* out_scalar;        // reduced value
    double* flag;              // T_NORMAL: #sigma<=0 written by a forward launch, read by a backward launch
    double* partials;          // [grid][kMaxCh]
    unsigned int* ticket;
    double* chan_adj[kMaxCh];  // adjoint destinations of broadcast scalar leaves
    int nch;
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void block_sum(Acc& a, double* sm)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int c = 0; c < kMaxCh; ++c) a.ch[c] = warp_sum(a.ch[c]);
    if (lane == 0)
#pragma unroll
        for (int c = 0; c < kMaxCh; ++c) sm[warp * kMaxCh + c] = a.ch[c];
    __syncthreads();
    if (warp == 0)
#pragma unroll
        for (int c = 0; c < kMaxCh; ++c) a.ch[c] = warp_sum(lane < nwarp ? sm[lane * kMaxCh + c] : 0.0);
    __syncthreads();
}

// sigma of the current element from the tape of a normal root (scalar sigma: same for every element)
template <class Root, class Tape>
__device__ __forceinline__ double last_tape_sigma(const Root&, const Tape&) { return 1.0; }
template <class X, class M, class S, class Tape>
__device__ __forceinline__ double last_tape_sigma(const Dev<core::NormalAdjLogPDFNode<X, M, S>>&, const Tape& t) { return t.ts.v; }
template <class L, class R, class Tape>
__device__ __forceinline__ double last_tape_sigma(const Dev<core::GlueNode<L, R>>& d, const Tape& t) { return last_tape_sigma(d.r, t.b); }

template <class Root>
__device__ __forceinline__ bool finish_normal(const Root&, Acc&, const Args&) { return false; }
template <class X, class M, class S>
__device__ __forceinline__ bool finish_normal(const Dev<core::NormalAdjLogPDFNode<X, M, S>>& nd, Acc& acc, const Args& a)
{
    double val;
    bool invalid;
    if (nd.sig_vec) {
        invalid = acc.ch[2] != 0.0;
        val = invalid ? -INFINITY : -0.5 * acc.ch[0] - acc.ch[1];                                      // normal.hpp:451, 548
    } else {
        typename Dev<S>::Tape ts;
        Acc dummy;
        nd.s.fwd(0, ts, dummy);                         // scalar sigma: every leaf below it is a broadcast
        const double sg = ts.v;
        invalid = !(sg > 0);
        if (invalid) val = -INFINITY;                                                                  // normal.hpp:147, 231, 344
        else if (nd.n == 1) val = -0.5 * acc.ch[0] - ::log(sg);
        else val = -0.5 * (acc.ch[0] / (sg * sg)) - (double)nd.n * ::log(sg);                            // normal.hpp:245-246
    }
    a.out_scalar[0] = val;
    if (a.flag) a.flag[0] = invalid ? 1.0 : 0.0;
    return invalid;
}
template <class L, class R>
__device__ __forceinline__ bool finish_normal(const Dev<core::GlueNode<L, R>>& d, Acc& acc, const Args& a)
{
    return finish_normal(d.r, acc, a);
}

template <class Root>
__global__ void __launch_bounds__(256)
fused_kernel(const __grid_constant__ Dev<Root> root, const __grid_constant__ Args a)
{
    constexpr int kind = terminal<Root>::kind;
    __shared__ double sm[32 * kMaxCh];
    __shared__ bool is_last;
    Acc acc;
#pragma unroll
    for (int c = 0; c < kMaxCh; ++c) acc.ch[c] = 0.0;
    bool back = a.mode & 2;
    bool gated = false;
    if (kind == T_NORMAL && (a.mode == 2) && a.flag) gated = *a.flag != 0.0;          // normal.hpp:457
    if (kind == T_NORMAL && !a.seed_vec && a.seed == 0.0) back = false;               // normal.hpp:160
    const unsigned long long tid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const unsigned long long nth = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = tid; i < a.N; i += nth) {
        typename Dev<Root>::Tape t;
        root.fwd(i, t, acc);
        if (kind == T_STORE && (a.mode & 1) && a.out_vec) a.out_vec[i] = t.v;
        if constexpr (kind == T_NORMAL) {
            if (!last_of(root).sig_vec && !(last_tape_sigma(root, t) > 0)) gated = true;               // normal.hpp:147
        }
        if (back && !gated) root.bwd(a.seed_vec ? a.seed_vec[i] : a.seed, i, t, acc);
    }
    if (kind == T_STORE && a.nch <= 3) return;
    // ---- deterministic grid reduction: CTA partials, last-arriving CTA sums them in CTA order
    block_sum(acc, sm);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int c = 0; c < kMaxCh; ++c) a.partials[(size_t)blockIdx.x * kMaxCh + c] = acc.ch[c];
        __threadfence();
        is_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
#pragma unroll
    for (int c = 0; c < kMaxCh; ++c) acc.ch[c] = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x)
#pragma unroll
        for (int c = 0; c < kMaxCh; ++c) acc.ch[c] += __ldcg(&a.partials[(size_t)b * kMaxCh + c]);
    block_sum(acc, sm);
    if (threadIdx.x != 0) return;
    *a.ticket = 0u;
    bool invalid = false;
    if (a.mode & 1) {
        if (kind == T_SUM) a.out_scalar[0] = acc.ch[0];
        if (kind == T_NORMAL) invalid = finish_normal(root, acc, a);
    }
    if (kind == T_NORMAL && a.mode == 2 && a.flag) invalid = *a.flag != 0.0;
    if ((a.mode & 2) && back && !invalid)
        for (int c = 3; c < a.nch; ++c) a.chan_adj[c][0] += acc.ch[c];                // var_view.hpp:91-94
}

}  // namespace fused

// --------------------------------------------------------------------------------- host API
namespace core {

template <class ExprType>
struct FusedBind {
    using expr_t = ExprType;
    using shape_t = typename ExprType::shape_t;
    static_assert(fused::root_fusable<ExprType>::value,
                  "ad::fuse: this expression contains a node the type-fused path does not cover "
                  "(dot, prod, sum/prod over expression lists, or a reduction that is not the root); use ad::bind");

    explicit FusedBind(const expr_t& expr, unsigned flags = 0) : expr_(expr), flags_(flags)
    {
        ctx_.N = domain_size(expr_);
        dev_ = fused::Dev<expr_t>::make(expr_, ctx_);
        size_ = expr_.size();
        const int threads = 256;
        int per_sm = 1, dev_id = 0, sms = 148;
        cudaGetDevice(&dev_id);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_id);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused::fused_kernel<expr_t>, threads, 0);
        size_t want = (ctx_.N + threads - 1) / threads;
        grid_ = (int)std::max<size_t>(1, std::min<size_t>(want, (size_t)sms * std::max(per_sm, 1)));
        void* p = nullptr;
        const size_t ws = ((size_t)grid_ * fused::kMaxCh + 4) * sizeof(double) + 64 + size_ * sizeof(double);
        detail::check(fadb_malloc(ws, &p));
        ws_.reset(static_cast<double*>(p), [](double* q) { fadb_free(q); });
        detail::check(fadb_memset_zero(p, ws));
        void* st = nullptr;
        detail::check(fadb_get_stream(&st));
        stream_ = static_cast<cudaStream_t>(st);
    }
    size_t size() const { return size_; }
    expr_t& get() { return expr_; }
    const double* value_device() const { return out_vec(); }

    // what: 1 forward, 2 backward, 3 both
    void run(int what, double seed, const double* seed_vec, double* host_value)
    {
        constexpr int kind = fused::terminal<expr_t>::kind;
        for (auto& s : ctx_.storages) s->to_device();
        fused::Args a{};
        a.N = ctx_.N; a.seed = seed; a.seed_vec = seed_vec;
        a.partials = ws_.get();
        a.out_scalar = ws_.get() + (size_t)grid_ * fused::kMaxCh;
        a.flag = a.out_scalar + 1;
        a.ticket = reinterpret_cast<unsigned int*>(a.out_scalar + 2);
        a.out_vec = kind == fused::T_STORE ? out_vec() : nullptr;
        a.nch = ctx_.next_ch;
        for (int c = 3; c < ctx_.next_ch; ++c) a.chan_adj[c] = ctx_.chan_adj[c - 3];
        // a variable vector sigma must be validated before any adjoint is written (normal.hpp:440-457)
        const bool two_pass = kind == fused::T_NORMAL && what == 3 && needs_sigma_check() && !(flags_ & FADB_TRUST_SIGMA);
        auto launch = [&](int mode) {
            a.mode = mode;
            fused::fused_kernel<expr_t><<<grid_, 256, 0, stream_>>>(dev_, a);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) throw Error(std::string("fastad_b200: fused kernel launch failed: ") + cudaGetErrorString(e));
        };
        if (two_pass) { launch(1); launch(2); }
        else launch(what);
        for (auto& s : ctx_.storages) {
            if (what & 1) s->dev_val_newer = true;
            if ((what & 2) && s->dadj) s->dev_adj_newer = true;
        }
        if (host_value && (what & 1)) {
            const double* src = kind == fused::T_STORE ? out_vec() : a.out_scalar;
            detail::check(fadb_d2h(host_value, src, size_ * sizeof(double)));
        }
    }

private:
    template <class E> static size_t domain_size(const E& e) { return e.size(); }
    template <class E> static size_t domain_size(const core::SumElemNode<E>& e) { return e.child().size(); }
    template <class E> static size_t domain_size(const core::NormNode<E>& e) { return e.child().size(); }
    template <class X, class M, class S>
    static size_t domain_size(const core::NormalAdjLogPDFNode<X, M, S>& e)
    {
        return std::max(e.x().size(), std::max(e.mean().size(), e.sigma().size()));
    }
    template <class L, class R> static size_t domain_size(const core::GlueNode<L, R>& e) { return domain_size(e.rhs()); }

    template <class E> static bool sigma_is_vec(const E&) { return false; }
    template <class X, class M, class S>
    static bool sigma_is_vec(const core::NormalAdjLogPDFNode<X, M, S>& e) { return e.sigma().size() > 1; }
    template <class L, class R> static bool sigma_is_vec(const core::GlueNode<L, R>& e) { return sigma_is_vec(e.rhs()); }
    bool needs_sigma_check() const { return sigma_is_vec(expr_); }
    double* out_vec() const { return ws_.get() + (size_t)grid_ * fused::kMaxCh + 4 + 8; }

    expr_t expr_;
    unsigned flags_;
    fused::Ctx ctx_;
    fused::Dev<expr_t> dev_;
    std::shared_ptr<double> ws_;
    cudaStream_t stream_ = nullptr;
    int grid_ = 1;
    size_t size_ = 1;
};

}  // namespace core

// ad::fuse(expr): the nvcc-side counterpart of ad::bind(expr)
template <class Derived>
inline auto fuse(const core::ExprBase<Derived>& expr, unsigned flags = 0)
{
    return core::FusedBind<Derived>(expr.self(), flags);
}

template <class E>
inline auto evaluate(core::FusedBind<E>& b)
{
    std::vector<double> v(b.size());
    b.run(1, 0., nullptr, v.data());
    return core::fetch(b, v);
}
template <class E, class = std::enable_if_t<util::is_scl_v<E>>>
inline void evaluate_adj(core::FusedBind<E>& b, double seed = 1.) { b.run(2, seed, nullptr, nullptr); }
template <class E>
inline void evaluate_adj(core::FusedBind<E>& b, device_seed seed) { b.run(2, 0., seed.ptr, nullptr); }
template <class E, class = std::enable_if_t<util::is_scl_v<E>>>
inline double autodiff(core::FusedBind<E>& b, double seed = 1.)
{
    double v = 0;
    b.run(3, seed, nullptr, &v);
    return v;
}
// vector-valued root: seed and value stay on the device (value_device())
template <class E>
inline void autodiff(core::FusedBind<E>& b, device_seed seed) { b.run(3, 0., seed.ptr, nullptr); }
// no host synchronisation at all (scalar root: the value stays at value_device() / is dropped)
template <class E>
inline void autodiff_async(core::FusedBind<E>& b, double seed = 1.) { b.run(3, seed, nullptr, nullptr); }

}  // namespace ad
