// expr.cu — generic expression path: builder, planner and the tape-interpreter kernels.
//
// This is the device side of ad::bind / ad::autodiff for ARBITRARY expressions built from the
// reference's hot-path nodes: VarView/Constant leaves (var_view.hpp, constant.hpp),
// UnaryNode (unary.hpp:32-120), BinaryNode (binary.hpp:36-185), PowNode (pow.hpp:67-196),
// SumElemNode/SumIterNode (sum.hpp), ProdElemNode/ProdIterNode (prod.hpp), DotNode (dot.hpp),
// NormalAdjLogPDFNode (stat/normal.hpp), EqNode (eq.hpp:29-159) and GlueNode (glue.hpp:25-121).
//
// The reference evaluates node by node, materialising every node's value and adjoint
// (bind.hpp:24-33).  Here the expression is cut into STAGES: a stage is a maximal region of
// same-shape elementwise nodes, optionally closed by a reduction (sum / prod / normal
// log-pdf); a DotNode is a stage of its own.  Inside a stage nothing is materialised: each
// thread interprets the stage's small SSA program for one element, forward then adjoint
// sweep, with node values in thread-local storage.  Only stage boundaries (a reduced scalar,
// a dot operand) and user-visible placeholders touch HBM.  The root stage runs forward and
// backward fused in one launch when its seed is known up front (sum.hpp:170-173 semantics);
// inner stages run forward, then backward (recomputing their forward values instead of
// caching them) once the consumer has produced their seed.
//
// Hot expression shapes are recognised at bind time and routed to the hand-written kernels
// (map_reduce.cu, normal.cu, linreg.cu); everything else runs on the interpreter below.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <array>
#include <atomic>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <vector>

#include "common.cuh"
#include "functors.cuh"
#include "jit.h"

extern "C" int fadb_sum_unary_async(int, size_t, const double*, double*, double, unsigned, double*);
extern "C" int fadb_sigma_check(size_t, const double*, double*);
extern "C" int fadb_normal_lpdf_partial(int, size_t, const double*, const double*, const double*, double*,
                                        double*, double*, double, unsigned, const double*, double*);
extern "C" int fadb_normal_lpdf_undo(int, size_t, const double*, const double*, const double*, double*, double*, double*, double,
                                     unsigned, const double*);
extern "C" int fadb_normal_lpdf_finish(int, size_t, const double*, const double*, const double*,
                                       const double*, double*, double*, double*, double, unsigned, double*);
extern "C" int fadb_linreg_partial(size_t, size_t, const double*, const double*, const double*, double*);
extern "C" int fadb_normal_dense_async(size_t, const double*, const double*, int, const double*, double*, double*, double*,
                                       double, unsigned, int, double*);
extern "C" int fadb_det_async(size_t, const double*, double*, double, unsigned, int, int, int, double*);
extern "C" int fadb_wishart_async(size_t, const double*, const double*, double, double*, double*, double, unsigned, int, double*);
extern "C" int fadb_logpdf_async(int, int, size_t, const double*, const double*, const double*, double*, double*, double*,
                                 double, unsigned, double*);
extern "C" int fadb_linreg_finish(size_t, size_t, const double*, const double*, double*, double*, double,
                                  unsigned, double*);
extern "C" int fadb_lsq_finish(size_t, const double*, double*, double, unsigned, double*);
extern "C" const char* fadb_jit_diagnostics(void);

#include "stage_kernel.cuh"

namespace fadb {

// tiled FP64 GEMM of dense_normal.cu: C = alpha op(A) op(B) + beta C, column-major
void gemm_f64(Context& c, bool ta, bool tb, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb,
              double beta, double* C, int ldc);

// ------------------------------------------------------------------------ DotNode kernels
// dot.hpp:89-104: V = L R; R_adj (+)= L^T A; L_adj (+)= A R^T.  Column-major, naive (one thread
// per output element): the large GEMV shape is routed to linreg.cu, these serve small operands.
__global__ void dot_fwd_kernel(int m, int k, int p, const double* __restrict__ L,
                               const double* __restrict__ R, double* V)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m * p) return;
    const int i = idx % m, j = idx / m;
    double acc = 0.0;
    for (int q = 0; q < k; ++q) acc += L[i + (size_t)q * m] * R[q + (size_t)j * k];
    V[idx] = acc;
}
__global__ void dot_bwd_kernel(int m, int k, int p, const double* __restrict__ L,
                               const double* __restrict__ R, const double* __restrict__ A,
                               double* Ladj, int lmode, double* Radj, int rmode)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (Radj && idx < k * p) {                     // (L^T A)(q, j)
        const int q = idx % k, j = idx / k;
        double acc = 0.0;
        for (int i = 0; i < m; ++i) acc += L[i + (size_t)q * m] * A[i + (size_t)j * m];
        Radj[idx] = (rmode == 2) ? acc : Radj[idx] + acc;
    }
    if (Ladj && idx < m * k) {                     // (A R^T)(i, q)
        const int i = idx % m, q = idx / m;
        double acc = 0.0;
        for (int j = 0; j < p; ++j) acc += A[i + (size_t)j * m] * R[q + (size_t)j * k];
        Ladj[idx] = (lmode == 2) ? acc : Ladj[idx] + acc;
    }
}
// TransposeNode (transpose.hpp:32-41): V = X^T; X_adj (+)= A^T.  X is m x k column-major.
// ---- tall operands: L is m x k with m large and R a k-vector (p == 1), the dot(X, w) of a statistical
// model.  HBM-bound: L crosses memory once per direction (8 m k bytes), coalesced down the columns.
constexpr int GV_THREADS = 256;
constexpr int GV_KT = 16;          // columns per register tile of the transposed product

// V = L R: KS threads share a row (thread (r, g) takes columns g, g + KS, ...; for a fixed column the rows
// of a CTA are contiguous, so every load is coalesced), partial dots meet in shared memory.  KS = 1 for very
// tall operands (one row per thread and trip), KS = 4 when there are too few rows to fill the GPU.
template <int KS>
__global__ void __launch_bounds__(GV_THREADS)
gemv_n_kernel(size_t m, int k, const double* __restrict__ L, const double* __restrict__ R, double* __restrict__ V)
{
    extern __shared__ double s_r[];                          // [k] (+ [KS][GV_THREADS / KS] partials when KS > 1)
    for (int q = threadIdx.x; q < k; q += blockDim.x) s_r[q] = R[q];
    __syncthreads();
    constexpr int RPB = GV_THREADS / KS;                     // rows per CTA and trip
    const int rl = threadIdx.x % RPB, g = threadIdx.x / RPB;
    double* s_p = s_r + k;
    for (size_t r0 = (size_t)blockIdx.x * RPB; r0 < m; r0 += (size_t)gridDim.x * RPB) {
        const size_t i = r0 + rl;
        double acc = 0.0;
        if (i < m) {
            const double* row = L + i;
#pragma unroll 8
            for (int q = g; q < k; q += KS) acc += __ldg(row + (size_t)q * m) * s_r[q];
        }
        if (KS == 1) {
            if (i < m) V[i] = acc;                           // dot.hpp:93, same order over q
        } else {
            s_p[g * RPB + rl] = acc;
            __syncthreads();
            if (g == 0 && i < m) {
                double v = s_p[rl];
#pragma unroll
                for (int t = 1; t < KS; ++t) v += s_p[t * RPB + rl];
                V[i] = v;
            }
            __syncthreads();
        }
    }
}

// Radj (+)= L^T A.  2D grid: blockIdx.x = block of rows, blockIdx.y = tile of GV_KT columns; a thread
// accumulates its rows' contribution to the GV_KT column sums in registers and the CTA reduces them ONCE
// (warp tree -> shared memory) into its partial row.  Deterministic: the last CTA to arrive adds the
// partial rows in row-block order.  L is read exactly once, A once per column tile (L2).
__global__ void __launch_bounds__(GV_THREADS)
gemv_t_kernel(size_t m, int k, const double* __restrict__ L, const double* __restrict__ A, double* partials /* [gridDim.x][k] */,
              unsigned int* ticket, double* Radj, int rmode)
{
    __shared__ double s_red[GV_THREADS / 32][GV_KT];
    __shared__ bool s_last;
    const size_t per = (m + gridDim.x - 1) / gridDim.x;
    const size_t r0 = (size_t)blockIdx.x * per, r1 = r0 + per < m ? r0 + per : m;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q0 = blockIdx.y * GV_KT;
    double acc[GV_KT];
#pragma unroll
    for (int t = 0; t < GV_KT; ++t) acc[t] = 0.0;
    if (q0 + GV_KT <= k) {
        for (size_t i = r0 + threadIdx.x; i < r1; i += blockDim.x) {
            const double a = __ldg(A + i);
#pragma unroll
            for (int t = 0; t < GV_KT; ++t) acc[t] += __ldg(L + i + (size_t)(q0 + t) * m) * a;
        }
    } else {
        for (size_t i = r0 + threadIdx.x; i < r1; i += blockDim.x) {
            const double a = __ldg(A + i);
#pragma unroll
            for (int t = 0; t < GV_KT; ++t)
                if (q0 + t < k) acc[t] += __ldg(L + i + (size_t)(q0 + t) * m) * a;
        }
    }
#pragma unroll
    for (int t = 0; t < GV_KT; ++t) {
        const double v = warp_sum(acc[t]);
        if (lane == 0) s_red[warp][t] = v;
    }
    __syncthreads();
    if (threadIdx.x < GV_KT && q0 + (int)threadIdx.x < k) {
        double v = 0.0;
        for (int w = 0; w < GV_THREADS / 32; ++w) v += s_red[w][threadIdx.x];
        partials[(size_t)blockIdx.x * k + q0 + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int q = threadIdx.x; q < k; q += blockDim.x) {
        double v = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) v += __ldcg(&partials[(size_t)b * k + q]);
        Radj[q] = (rmode == 2) ? v : Radj[q] + v;
    }
    if (threadIdx.x == 0) *ticket = 0u;
}

// Ladj (+)= A R^T for p == 1: an outer product, one element per thread (only when L is a variable)
__global__ void outer_kernel(size_t m, int k, const double* __restrict__ A, const double* __restrict__ R, double* Ladj, int lmode)
{
    const size_t n = m * (size_t)k, nth = (size_t)gridDim.x * blockDim.x;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += nth) {
        const double v = A[idx % m] * R[idx / m];
        Ladj[idx] = (lmode == 2) ? v : Ladj[idx] + v;
    }
}

// Ladj (+)= A R^T when the SHARED dimension p is the large one (a small weight matrix times a batch:
// dot(W, X) with X k x p, p = batch): the m x k result is tiny, the sum runs over p columns.  Every CTA
// takes a contiguous range of columns, thread e owns output entry e (m k <= GV_THREADS * WR_PER), per-CTA
// partial results, last CTA adds them in CTA order (deterministic).  A and R are read exactly once.
constexpr int WR_PER = 4;
__global__ void __launch_bounds__(GV_THREADS)
wide_outer_kernel(int m, int k, size_t p, const double* __restrict__ A, const double* __restrict__ R, double* partials /* [grid][m k] */,
                  unsigned int* ticket, double* Ladj, int lmode)
{
    __shared__ bool s_last;
    const int n = m * k;
    const size_t per = (p + gridDim.x - 1) / gridDim.x;
    const size_t j0 = (size_t)blockIdx.x * per, j1 = j0 + per < p ? j0 + per : p;
    double acc[WR_PER];
    int ei[WR_PER], eq[WR_PER];
#pragma unroll
    for (int t = 0; t < WR_PER; ++t) {
        const int e = threadIdx.x + t * GV_THREADS;
        acc[t] = 0.0; ei[t] = e < n ? e % m : 0; eq[t] = e < n ? e / m : 0;
    }
    for (size_t j = j0; j < j1; ++j) {
        const double* a = A + j * (size_t)m;
        const double* r = R + j * (size_t)k;
#pragma unroll
        for (int t = 0; t < WR_PER; ++t) acc[t] += __ldg(a + ei[t]) * __ldg(r + eq[t]);
    }
#pragma unroll
    for (int t = 0; t < WR_PER; ++t) {
        const int e = threadIdx.x + t * GV_THREADS;
        if (e < n) partials[(size_t)blockIdx.x * n + e] = acc[t];
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        double v = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) v += __ldcg(&partials[(size_t)b * n + e]);
        Ladj[e] = (lmode == 2) ? v : Ladj[e] + v;
    }
    if (threadIdx.x == 0) *ticket = 0u;
}

// The same sum for a SMALL result (m k <= WS_MAXN, e.g. a 3 x 2 weight matrix): a thread per output entry
// would leave almost every thread idle, so the columns are spread over the lanes instead: a warp takes 32
// consecutive columns (coalesced), each product is summed across the warp and lane 0 adds it to the warp's
// row of shared-memory accumulators; warps, CTAs and the final pass combine in a fixed order.
constexpr int WS_MAXN = 256;
__global__ void __launch_bounds__(GV_THREADS)
wide_outer_small_kernel(int m, int k, size_t p, const double* __restrict__ A, const double* __restrict__ R, double* partials,
                        unsigned int* ticket, double* Ladj, int lmode)
{
    __shared__ double s_acc[GV_THREADS / 32][WS_MAXN];
    __shared__ bool s_last;
    const int n = m * k, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = GV_THREADS / 32;
    for (int e = lane; e < n; e += 32) s_acc[warp][e] = 0.0;
    __syncwarp();
    const size_t per = (p + gridDim.x - 1) / gridDim.x;
    const size_t j0 = (size_t)blockIdx.x * per, j1 = j0 + per < p ? j0 + per : p;
    for (size_t jb = j0 + (size_t)warp * 32; jb < j1; jb += (size_t)nw * 32) {
        const size_t j = jb + lane;
        const bool in = j < j1;
        for (int q = 0; q < k; ++q) {
            const double rq = in ? __ldg(R + j * (size_t)k + q) : 0.0;
            for (int i = 0; i < m; ++i) {
                double v = in ? __ldg(A + j * (size_t)m + i) * rq : 0.0;
                v = warp_sum(v);
                if (lane == 0) s_acc[warp][i + q * m] += v;
            }
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        double v = 0.0;
        for (int w = 0; w < nw; ++w) v += s_acc[w][e];
        partials[(size_t)blockIdx.x * n + e] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        double v = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) v += __ldcg(&partials[(size_t)b * n + e]);
        Ladj[e] = (lmode == 2) ? v : Ladj[e] + v;
    }
    if (threadIdx.x == 0) *ticket = 0u;
}

__global__ void transpose_fwd_kernel(int m, int k, const double* __restrict__ X, double* V)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m * k) return;
    const int i = idx % m, j = idx / m;
    V[j + (size_t)i * k] = X[idx];
}
__global__ void transpose_bwd_kernel(int m, int k, const double* __restrict__ A, double* Xadj, int mode)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m * k) return;
    const int i = idx % m, j = idx / m;
    const double a = A[j + (size_t)i * k];
    Xadj[idx] = (mode == 2) ? a : Xadj[idx] + a;
}
// ---- one-pass dot -> map -> sum (csrc/glm_kernel.cuh, compiled by NVRTC): host mirror of its argument block and the
// scalar epilogue.  out = {channel totals [nacc], X^T a [cols]} as the kernel's last CTA left them.
struct GlmArgsHost {
    unsigned long long rows;
    int cols, pad;
    const double* w;
    double seed;
    double* cta_partials;
    unsigned int* ticket;
    double* out;
    const double* lp_val[MAXL_JIT];
};
struct GlmFinish {
    int nacc, cols, nscal, term;
    double n_total; const double* sig_ptr;   // T_NORMAL: the one sigma (device) and the row count
    double* w_adj; int w_mode;           // 0 none, 1 accumulate, 2 assign
    double* s_adj[MAXCH_JIT]; int s_ch[MAXCH_JIT]; int s_mode[MAXCH_JIT];
};
__global__ void glm_finish_kernel(GlmFinish f, const double* __restrict__ out, double* value)
{
    const int t = threadIdx.x;
    if (f.term == T_NORMAL) {
        const double s = f.sig_ptr[0];
        if (!(s > 0)) { if (t == 0) value[0] = -INFINITY; return; }         // normal.hpp:344: no adjoint update either
        if (t == 0) value[0] = -0.5 * (out[0] / (s * s)) - f.n_total * log(s);   // normal.hpp:351-352
    } else if (t == 0) value[0] = out[0];                                    // SumElemNode::feval (sum.hpp:156-160)
    if (t < f.nscal) {                                                       // VarView::beval of a scalar variable (var_view.hpp:91-94)
        const double g = out[f.s_ch[t]];
        f.s_adj[t][0] = f.s_mode[t] == 2 ? g : f.s_adj[t][0] + g;
    }
    if (f.w_adj)
        for (int j = t; j < f.cols; j += blockDim.x) {                       // DotNode::beval: w_adj += X^T a (dot.hpp:100)
            const double g = out[f.nacc + j];
            f.w_adj[j] = f.w_mode == 2 ? g : f.w_adj[j] + g;
        }
}
int linreg_encode_xmap(void* map128, const double* X, size_t rows, size_t cols);   // linreg.cu (CUtensorMap, 128 bytes, 64-aligned)

__global__ void fill_kernel(double* p, size_t n, double v)
{
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ======================================================================== host: builder + planner
enum NodeKind { N_VAR, N_CONST_S, N_CONST_V, N_UNARY, N_BINARY, N_POW, N_SUM, N_SUM_ITER, N_PROD,
                N_PROD_ITER, N_DOT, N_NORMAL, N_EQ, N_GLUE, N_TRANSPOSE, N_IF_ELSE, N_OPEQ, N_LOGPDF, N_DET, N_WISHART };

struct ENode {
    NodeKind kind;
    int fn = 0;
    long pw = 0;
    int shape = FADB_SCL;
    size_t rows = 1, cols = 1;
    std::vector<int> ch;
    const double* val = nullptr;   // leaves: device pointers
    double* adj = nullptr;
    double cval = 0;
    size_t size() const { return rows * cols; }
};

struct OperandRef { const double* val = nullptr; double* adj = nullptr; int adj_mode = 0; int node = -1; int stage = -1; };

inline unsigned long long next_stage_uid() { static std::atomic<unsigned long long> n{0}; return ++n; }

struct Stage {
    unsigned long long uid = next_stage_uid();
    int type = 0;                  // 0 = MAP (interpreter), 1 = DOT, 2 = TRANSPOSE, 3 = dense-covariance normal,
                                   // 4 = det / log_det (k = take_log, p = policy), 5 = wishart (sig_const = df)
    size_t N = 1;
    std::vector<Ins> prog;
    std::vector<LeafDesc> leaves;
    std::vector<int> leaf_node;    // node id of each leaf (-1: stage boundary)
    std::vector<int> leaf_stage;   // producing stage of a boundary leaf (-1 otherwise)
    std::map<int, int> eq_slot;    // placeholder node id -> slot holding its current value
    std::set<int> assigned;        // placeholder node ids assigned in this stage
    std::map<int, int> boundary_slot;   // node id of a producer-stage result -> the slot that already reads it in this stage
    int term = T_NONE, root = -1, nx = -1, nm = -1, ns = -1, mu_vec = 0, sig_vec = 0;
    int sig_leaf = -1;             // leaf index of a scalar sigma (T_NORMAL)
    bool needs_check = false;      // log-pdf whose validity is a property of ALL elements
    double sig_const = 0; bool sig_is_const = false;
    int nch = 0, term_ch = 0;
    int node = -1;                 // expression node this stage evaluates (its output)
    size_t out_size = 1;
    // DOT
    int m = 0, k = 0, p = 0;
    OperandRef L, R, S3;           // S3: Sigma of a dense-covariance normal (L = x, R = mu)
    bool sigma_const = false, factored = false;
    // device state
    double* mval = nullptr;        // materialised output value [out_size]
    double* madj = nullptr;        // materialised output adjoint [out_size] (inner stages)
    DIns* d_prog = nullptr;
    std::vector<DIns> enc;         // device encoding of `prog` (also the text of the specialised kernel)
    JitKernel jit[4];              // specialised kernels by mode (M_F, M_B, M_FB), compiled on first use
    int jit_state[4] = {0, 0, 0, 0};   // 0 untried, 1 ready, -1 not specialisable / compile failed
    JitKernel jit_ring[4];         // the same with the cp.async operand ring (stage_kernel.cuh, JIT_RING)
    int jit_ring_state[4] = {0, 0, 0, 0};
    LeafDesc* d_leaves = nullptr;
    // Leaves that this stage READS from memory and later ASSIGNS (w op= expr, or a read of w ahead of `w = expr`):
    // a backward-only launch (M_B) recomputes the forward values, but the forward-only launch (M_F) has already
    // overwritten w.  Like OpEqNode's cache_ (eq.hpp:240-271) the pre-stage value is kept: M_F saves it, M_B reads
    // it instead of w.val and, for op=, restores w from it afterwards.
    struct PreVal { int leaf; double* cache; size_t n; bool restore; };
    std::vector<PreVal> pre;
    LeafDesc* d_leaves_b = nullptr;    // d_leaves with those leaves' val redirected to their cache (M_B launches)
    double* d_red = nullptr;       // [8]
    double* d_sig = nullptr;       // device copy of a constant scalar sigma
    bool big = false;              // scalar stage too large for the on-chip interpreter
    bool jit_only = false;         // vector stage beyond the interpreter's limits: specialised kernel only
    double* d_scratch = nullptr;
    int fast = 0;                  // 0 interpreter, 1 sum_unary, 2 normal leaves, 3 linreg, 4 cauchy/uniform/bernoulli leaves
    // fast-path operands
    int f_op = 0; const double* f_x = nullptr; double* f_xadj = nullptr;
    const double *f_mu = nullptr, *f_sig = nullptr; double *f_muadj = nullptr, *f_sigadj = nullptr;
    int f_shape = 0; const double* f_X = nullptr; const double* f_w = nullptr; double* f_wadj = nullptr;
    size_t f_rows = 0, f_cols = 0; double* f_part = nullptr;
    // fast == 6: the one-pass dot -> map -> sum kernel (glm_kernel.cuh): leaf of the dot result, per-row operand slots
    int g_dot_leaf = -1, g_nvec = 0; std::vector<int> g_vec_slot; int g_dot_stage = -1;
    JitKernel g_jit; int g_jit_state = 0; double* g_ws = nullptr; size_t g_ws_cap = 0;
    double* gv_part = nullptr; int gv_grid = 0;    // tall dot stage: per-CTA partial rows of L^T A
    double* wr_part = nullptr; int wr_grid = 0;    // wide dot stage (large p, small m x k): per-CTA partials of A R^T
};

}  // namespace fadb

using namespace fadb;

static std::atomic<int> g_glm_enabled{1};     // fadb_glm_enable: the one-pass dot -> map -> sum route (A/B runs, tests)
static std::atomic<int> g_ring_enabled{1};    // fadb_jit_ring_enable: the cp.async operand ring of placeholder stages (A/B runs, tests)
static std::atomic<int> g_additive_enabled{1};   // fadb_additive_root_enable: per-term execution of additive roots

struct fadb_expr {
    std::vector<ENode> nodes;
    std::vector<std::unique_ptr<Stage>> plan;
    std::vector<void*> owned;        // device allocations owned by the expression
    int root = -1;
    bool planned = false, bound = false;
    unsigned flags = 0;
    std::string err;
    size_t cache_vals = 0, cache_adjs = 0;
    std::map<std::array<long, 5>, int> pure_index;     // (kind, functor, exponent, child, child) -> node id, see find_pure()
    size_t pure_indexed = 0;

    int add(ENode&& n) { nodes.push_back(std::move(n)); return (int)nodes.size() - 1; }
    bool ok(int id) const { return id >= 0 && id < (int)nodes.size(); }
    bool is_const_scalar(int id) const { return nodes[id].kind == N_CONST_S; }

    // ---------------------------------------------------------------- planning (host only)
    static bool barrier(NodeKind k) { return k == N_SUM || k == N_PROD || k == N_NORMAL || k == N_DOT || k == N_TRANSPOSE || k == N_LOGPDF ||
                                                 k == N_DET || k == N_WISHART; }

    int emit(Stage& st, Ins I)
    {
        st.prog.push_back(I);
        return (int)st.prog.size() - 1;
    }
    int leaf_index(Stage& st, int nid, const double* val, double* adj, bool scalar, int adj_mode, int from_stage)
    {
        if (nid >= 0)
            for (size_t k = 0; k < st.leaf_node.size(); ++k)
                if (st.leaf_node[k] == nid) return (int)k;
        LeafDesc L{val, adj, scalar ? 1 : 0, adj ? adj_mode : 0, -1, 0};
        st.leaves.push_back(L);
        st.leaf_node.push_back(nid);
        st.leaf_stage.push_back(from_stage);
        return (int)st.leaves.size() - 1;
    }

    // value of `nid` as a slot of stage `st` (emits instructions, creates producer stages)
    int lower(int nid, Stage& st)
    {
        ENode& n = nodes[nid];
        const bool same = n.size() == st.N;
        const bool bcast = n.size() == 1 && st.N > 1;
        switch (n.kind) {
        case N_VAR: {
            auto it = st.eq_slot.find(nid);
            if (it != st.eq_slot.end()) return it->second;
            if (!same && !bcast) { err = "shape mismatch between a leaf and its elementwise consumer"; return -1; }
            int lf = leaf_index(st, nid, n.val, n.adj, bcast, 1, -1);
            return emit(st, Ins{I_LEAF, 0, 0, 0, lf, 0, 0.0});
        }
        case N_CONST_S: return emit(st, Ins{I_CONST, 0, 0, 0, 0, 0, n.cval});
        case N_CONST_V: {
            if (!same && !bcast) { err = "shape mismatch between a constant view and its consumer"; return -1; }
            int lf = leaf_index(st, nid, n.val, nullptr, bcast, 0, -1);
            return emit(st, Ins{I_LEAF, 0, 0, 0, lf, 0, 0.0});
        }
        default: break;
        }
        if (barrier(n.kind) || !same) {
            // produced by another stage: a reduction, a dot, or a sub-expression of another shape
            if (!barrier(n.kind) && !bcast) { err = "shape mismatch between elementwise operands"; return -1; }
            // The same producer node read twice by this stage (eta = dot(X, w) + b used in two terms): the reference evaluates
            // its copy of the sub-expression twice with identical results; here it is planned once and its slot reused, the
            // adjoints of both uses meet in that slot.  Not across an assignment of this stage (the operands may have changed).
            auto seen = st.boundary_slot.find(nid);
            if (seen != st.boundary_slot.end()) return seen->second;
            int sidx = plan_stage(nid);
            if (sidx < 0) return -1;
            Stage& ps = *plan[sidx];
            const bool sc = ps.out_size == 1 && st.N > 1;
            if (ps.out_size != st.N && !sc) { err = "shape mismatch at a stage boundary"; return -1; }
            int lf = leaf_index(st, -1, nullptr, nullptr, sc, 2, sidx);
            st.leaves[lf].adj_mode = 2;       // pointers are patched at bind time
            const int slot = emit(st, Ins{I_LEAF, 0, 0, 0, lf, 0, 0.0});
            st.boundary_slot[nid] = slot;
            return slot;
        }
        switch (n.kind) {
        case N_UNARY: {
            int a = lower(n.ch[0], st); if (a < 0) return -1;
            return emit(st, Ins{I_UNARY, n.fn, a, 0, 0, 0, 0.0});
        }
        case N_POW: {
            int a = lower(n.ch[0], st); if (a < 0) return -1;
            return emit(st, Ins{I_POW, (int)n.pw, a, 0, 0, 0, 0.0});
        }
        case N_BINARY: {
            int a = lower(n.ch[0], st); if (a < 0) return -1;      // lhs first (binary.hpp:102-103)
            int b = lower(n.ch[1], st); if (b < 0) return -1;
            return emit(st, Ins{I_BINARY, n.fn, a, b, 0, 0, 0.0});
        }
        case N_SUM_ITER: case N_PROD_ITER: {
            // sum.hpp:57-64 / prod.hpp:55-62: left-to-right accumulation over the sub-expressions
            if (n.ch.empty()) return emit(st, Ins{I_CONST, 0, 0, 0, 0, 0, n.kind == N_SUM_ITER ? 0.0 : 1.0});
            int acc = lower(n.ch[0], st); if (acc < 0) return -1;
            for (size_t k = 1; k < n.ch.size(); ++k) {
                int b = lower(n.ch[k], st); if (b < 0) return -1;
                acc = emit(st, Ins{I_BINARY, n.kind == N_SUM_ITER ? FADB_ADD : FADB_MUL, acc, b, 0, 0, 0.0});
            }
            return acc;
        }
        case N_EQ: {
            ENode& w = nodes[n.ch[0]];
            int a = lower(n.ch[1], st); if (a < 0) return -1;
            int lf = leaf_index(st, n.ch[0], w.val, w.adj, false, 1, -1);
            st.leaves[lf].assigned = 1;
            int s = emit(st, Ins{I_EQ, 0, a, 0, lf, 0, 0.0});
            st.eq_slot[n.ch[0]] = s;
            st.assigned.insert(n.ch[0]);
            st.boundary_slot.clear();
            return s;
        }
        case N_GLUE: {
            int a = lower(n.ch[0], st); if (a < 0) return -1;
            int b = lower(n.ch[1], st); if (b < 0) return -1;
            return emit(st, Ins{I_GLUE, 0, a, b, 0, 0, 0.0});
        }
        case N_IF_ELSE: {
            // cond ; JZ else ; <if> ; JMP end ; else: <else> ; end: PHI    (if_else.hpp:64-78)
            int c = lower(n.ch[0], st); if (c < 0) return -1;
            int jz = emit(st, Ins{I_JZ, 0, c, 0, 0, 0, 0.0});
            const std::map<int, int> outer = st.boundary_slot;   // slots loaded inside a branch exist only when it is taken
            int a = lower(n.ch[1], st); if (a < 0) return -1;
            int jmp = emit(st, Ins{I_JMP, 0, c, 0, jz, 0, 0.0});
            st.prog[jz].b = (int)st.prog.size();
            st.boundary_slot = outer;
            int b = lower(n.ch[2], st); if (b < 0) return -1;
            st.prog[jmp].b = (int)st.prog.size();
            st.boundary_slot = outer;
            return emit(st, Ins{I_PHI, 0, a, b, c, 0, (double)jmp});
        }
        case N_OPEQ: {
            ENode& w = nodes[n.ch[0]];
            int old = lower(n.ch[0], st); if (old < 0) return -1;      // previous value (leaf read or alias slot)
            int a = lower(n.ch[1], st); if (a < 0) return -1;
            int lf = leaf_index(st, n.ch[0], w.val, w.adj, false, 1, -1);
            st.leaves[lf].assigned = 1;
            int s = emit(st, Ins{I_OPEQ, n.fn, a, old, lf, 0, 0.0});
            st.eq_slot[n.ch[0]] = s;
            st.assigned.insert(n.ch[0]);
            st.boundary_slot.clear();
            return s;
        }
        default: break;
        }
        err = "internal: unhandled node kind in lower()";
        return -1;
    }

    OperandRef operand(int nid)
    {
        ENode& n = nodes[nid];
        OperandRef r;
        r.node = nid;
        if (n.kind == N_VAR) { r.val = n.val; r.adj = n.adj; r.adj_mode = n.adj ? 1 : 0; return r; }
        if (n.kind == N_CONST_V) { r.val = n.val; return r; }
        r.stage = plan_stage(nid);
        r.adj_mode = 2;
        return r;
    }

    bool finish_stage(Stage& st)
    {
        st.big = (int)st.prog.size() > MAXS || (int)st.leaves.size() > MAXL;
        if (st.big && st.N > 1) {
            // too large for the interpreter's fixed arrays: fine for a specialised kernel (if_else included: jump watermark)
            if ((int)st.prog.size() > MAXS_JIT || (int)st.leaves.size() > MAXL_JIT) {
                err = "a fused vector stage is limited to " + std::to_string(MAXS_JIT) + " nodes and " + std::to_string(MAXL_JIT) + " distinct leaves";
                return false;
            }
            st.big = false;
            st.jit_only = true;
        }
        st.term_ch = st.term == T_SUM ? 1 : st.term >= T_NORMAL ? 3 : st.term == T_PROD ? 1 : 0;
        int ch = st.term_ch;
        for (auto& L : st.leaves)
            if (L.scalar && L.adj_mode != 0 && st.N > 1) L.channel = ch++;
        if (ch > MAXCH_JIT) { err = "expression stage needs more than " + std::to_string(MAXCH_JIT) + " reduction channels"; return false; }
        if (ch > MAXCH) st.jit_only = true;
        st.nch = ch;
        return true;
    }

    // plans the stage that produces node `nid`; returns its index in `plan`
    int plan_stage(int nid)
    {
        ENode& n = nodes[nid];
        auto st = std::make_unique<Stage>();
        st->node = nid;
        if (n.kind == N_DOT) {
            ENode& l = nodes[n.ch[0]];
            ENode& r = nodes[n.ch[1]];
            if (l.cols != r.rows) { err = "dot: inner dimensions differ"; return -1; }
            st->type = 1;
            // rows of a tall L are indexed with size_t; every other extent and every small-operand product is an int
            constexpr size_t IMAX = 0x7fffffffu;
            if (l.rows > IMAX || l.cols > IMAX || r.cols > IMAX || l.cols * r.cols > IMAX) {
                err = "dot: an extent (or the right operand) exceeds 2^31 - 1";
                return -1;
            }
            st->m = (int)l.rows; st->k = (int)l.cols; st->p = (int)r.cols;
            st->out_size = n.size();
            st->L = operand(n.ch[0]); if (!err.empty()) return -1;
            st->R = operand(n.ch[1]); if (!err.empty()) return -1;
            const bool tall = r.cols <= 16 && l.rows >= 4096 && l.cols <= 4096;      // the size_t-indexed GEMV kernels
            if ((l.rows * l.cols > IMAX || l.rows * r.cols > IMAX) && !(tall && (r.cols == 1 || !st->L.adj_mode))) {
                err = "dot: operand exceeds 2^31 - 1 elements and is not a tall matrix times a vector";
                return -1;
            }
        } else if (n.kind == N_TRANSPOSE) {
            ENode& x = nodes[n.ch[0]];
            st->type = 2;
            if (x.rows * x.cols > 0x7fffffffu) { err = "transpose: operand exceeds 2^31 - 1 elements"; return -1; }
            st->m = (int)x.rows; st->k = (int)x.cols; st->p = 0;
            st->out_size = n.size();
            st->L = operand(n.ch[0]); if (!err.empty()) return -1;
        } else if (n.kind == N_DET) {
            // DetNode / LogDetNode (det.hpp:25-92): the matrix is handed to the dense kernels as a whole operand
            st->type = 4;
            st->m = (int)nodes[n.ch[0]].rows; st->k = (int)n.pw; st->p = n.fn;
            st->out_size = 1;
            st->L = operand(n.ch[0]); if (!err.empty()) return -1;
        } else if (n.kind == N_WISHART) {
            st->type = 5;
            st->m = (int)nodes[n.ch[0]].rows;
            st->sig_const = n.cval;
            st->out_size = 1;
            st->L = operand(n.ch[0]); if (!err.empty()) return -1;
            st->R = operand(n.ch[1]); if (!err.empty()) return -1;
        } else if (n.kind == N_SUM || n.kind == N_PROD) {
            st->N = nodes[n.ch[0]].size();
            st->term = n.kind == N_SUM ? T_SUM : T_PROD;
            st->root = lower(n.ch[0], *st);
            if (st->root < 0) return -1;
        } else if (n.kind == N_LOGPDF) {
            size_t N = 1;
            for (int c : n.ch) N = std::max(N, nodes[c].size());
            st->N = N;
            st->term = T_CAUCHY + n.fn;
            st->nx = lower(n.ch[0], *st); if (st->nx < 0) return -1;
            st->nm = lower(n.ch[1], *st); if (st->nm < 0) return -1;
            if (n.ch.size() > 2) { st->ns = lower(n.ch[2], *st); if (st->ns < 0) return -1; }
            st->sig_vec = n.ch.size() > 2 && nodes[n.ch[2]].size() > 1;
            // Cauchy with a scalar scale is the only case whose validity every thread can see
            st->needs_check = !(st->term == T_CAUCHY && !st->sig_vec);
        } else if (n.kind == N_NORMAL && nodes[n.ch[2]].shape == FADB_MAT && nodes[n.ch[2]].rows == nodes[n.ch[2]].cols &&
                   nodes[n.ch[2]].rows == nodes[n.ch[0]].size() && nodes[n.ch[2]].rows > 1) {
            // vsm / vvm (normal.hpp:581-773): x, mu, Sigma are handed to K7 as whole operands
            st->type = 3;
            st->m = (int)nodes[n.ch[0]].size();
            st->k = nodes[n.ch[1]].size() > 1 ? 1 : 0;          // mu is a vector
            st->out_size = 1;
            st->L = operand(n.ch[0]); if (!err.empty()) return -1;
            st->R = operand(n.ch[1]); if (!err.empty()) return -1;
            st->S3 = operand(n.ch[2]); if (!err.empty()) return -1;
            st->sigma_const = nodes[n.ch[2]].kind == N_CONST_V;
        } else if (n.kind == N_NORMAL) {
            size_t N = 1;
            for (int c : n.ch) N = std::max(N, nodes[c].size());
            st->N = N;
            st->term = T_NORMAL;
            st->nx = lower(n.ch[0], *st); if (st->nx < 0) return -1;
            st->nm = lower(n.ch[1], *st); if (st->nm < 0) return -1;
            st->ns = lower(n.ch[2], *st); if (st->ns < 0) return -1;
            st->mu_vec = nodes[n.ch[1]].size() > 1;
            st->sig_vec = nodes[n.ch[2]].size() > 1;
            if (!st->sig_vec) {
                const Ins& I = st->prog[st->ns];
                if (I.op == I_LEAF) st->sig_leaf = I.leaf;
                else if (I.op == I_CONST) { st->sig_is_const = true; st->sig_const = I.imm; }
                else if (N > 1) { err = "internal: scalar sigma must be a leaf of its stage"; return -1; }
                else {   // scalar stage: sigma is computed in-stage; materialise through a producer
                    err = "normal_adj_log_pdf: scalar sigma expression inside a scalar stage is not supported yet";
                    return -1;
                }
            }
        } else {
            st->N = n.size();
            st->out_size = n.size();
            st->root = lower(nid, *st);
            if (st->root < 0) return -1;
        }
        if (st->type == 0 && !finish_stage(*st)) return -1;
        plan.push_back(std::move(st));
        return (int)plan.size() - 1;
    }

    void flatten(int nid, std::vector<int>& out)
    {
        if (nodes[nid].kind == N_GLUE) { flatten(nodes[nid].ch[0], out); flatten(nodes[nid].ch[1], out); }
        else out.push_back(nid);
    }
    // does `nid` read one of `assigned` from behind a stage boundary (relative to domain N)?
    bool reads_across(int nid, size_t N, const std::set<int>& assigned, bool crossed)
    {
        ENode& n = nodes[nid];
        if (n.kind == N_VAR) return crossed && assigned.count(nid);
        bool c2 = crossed || barrier(n.kind) || n.size() != N;
        for (size_t k = 0; k < n.ch.size(); ++k) {
            if (n.kind == N_EQ && k == 0) continue;
            if (reads_across(n.ch[k], N, assigned, c2)) return true;
        }
        return false;
    }

    int do_plan(int root_id)
    {
        plan.clear(); err.clear();
        root = root_id;
        std::vector<int> stmts;
        flatten(root_id, stmts);
        std::unique_ptr<Stage> open;
        auto close = [&]() -> bool {
            if (!open) return true;
            if (!finish_stage(*open)) return false;
            plan.push_back(std::move(open));
            return true;
        };
        for (int s : stmts) {
            ENode& n = nodes[s];
            if (barrier(n.kind)) {
                if (!close()) return FADB_ERR_UNSUPPORTED;
                if (plan_stage(s) < 0) return FADB_ERR_UNSUPPORTED;
                continue;
            }
            if (open && open->N == n.size() && !reads_across(s, open->N, open->assigned, false)) {
                int slot = lower(s, *open);
                if (slot < 0) return FADB_ERR_UNSUPPORTED;
                open->root = slot; open->node = s;
                continue;
            }
            if (!close()) return FADB_ERR_UNSUPPORTED;
            open = std::make_unique<Stage>();
            open->N = n.size(); open->out_size = n.size(); open->node = s;
            open->root = lower(s, *open);
            if (open->root < 0) return FADB_ERR_UNSUPPORTED;
        }
        if (!close()) return FADB_ERR_UNSUPPORTED;
        detect_fast_paths();
        // reference-compatible cache accounting: only stage boundaries are materialised
        cache_vals = cache_adjs = 0;
        for (size_t k = 0; k < plan.size(); ++k) {
            cache_vals += plan[k]->out_size;
            if (k + 1 < plan.size()) cache_adjs += plan[k]->out_size;
        }
        planned = true;
        return FADB_OK;
    }

    bool leaf_only(const Stage& st, int slot, int* leaf_out) const
    {
        const Ins& I = st.prog[slot];
        if (I.op != I_LEAF) return false;
        if (st.leaf_node[I.leaf] < 0) return false;
        *leaf_out = I.leaf;
        return true;
    }

    bool detect_lsq(Stage& st)
    {
        if (st.prog.size() != 4 || st.root != 3) return false;
        const Ins &Ip = st.prog[3], &Is = st.prog[2];
        if (Ip.op != I_POW || Ip.fn != 2 || Ip.a != 2 || Is.op != I_BINARY || Is.fn != FADB_SUB) return false;
        if (st.prog[Is.a].op != I_LEAF || st.prog[Is.b].op != I_LEAF) return false;
        int ld = st.prog[Is.a].leaf, ly = st.prog[Is.b].leaf;
        if (st.leaf_stage[ld] < 0) std::swap(ld, ly);
        if (st.leaf_stage[ld] < 0 || st.leaf_stage[ly] >= 0 || st.leaves[ly].scalar || st.leaves[ly].adj_mode != 0) return false;
        Stage& ds = *plan[st.leaf_stage[ld]];
        if (ds.type != 1 || ds.p != 1 || ds.L.stage >= 0 || ds.R.stage >= 0 || ds.L.adj_mode != 0 || ds.k > 64) return false;   // beyond one 64-column block the two-pass GEMV kernels are faster (measured)
        st.fast = 5;
        st.f_X = ds.L.val; st.f_w = ds.R.val; st.f_wadj = ds.R.adj_mode ? ds.R.adj : nullptr;
        st.f_rows = ds.m; st.f_cols = ds.k;
        st.f_x = st.leaves[ly].val;
        return true;
    }

    // sum(f(dot(X, w), per-row constants, scalars)) with X constant and at most 64 columns: the row adjoint of the product
    // depends on the row alone (root seed known up front), so the map runs inside the one-pass TMA kernel (glm_kernel.cuh)
    // and X crosses HBM once instead of twice.  Anything the kernel does not cover keeps fast == 0 (gemv -> stage -> gemv^T).
    void detect_glm(Stage& st)
    {
        static const bool off = getenv("FADB_NO_GLM") != nullptr;
        if (off || st.big || st.prog.size() > 48 || st.nch > 16 || st.N < 64 || st.N % 2 != 0) return;     // the TMA tile path needs an even row count
        if (st.term == T_NORMAL && (st.sig_vec || !st.mu_vec || (!st.sig_is_const && st.sig_leaf < 0))) return;
        int ld = -1, nvec = 0;
        std::vector<int> slot(st.leaves.size(), -1);
        for (size_t l = 0; l < st.leaves.size(); ++l) {
            const LeafDesc& L = st.leaves[l];
            if (st.leaf_stage[l] >= 0) { if (ld >= 0 || L.scalar) return; ld = (int)l; continue; }
            if (L.assigned) return;
            if (L.scalar) { if (L.adj_mode == 2) return; continue; }
            if (L.adj_mode != 0 || nvec >= 4) return;         // per-row operands: constants only (y, offsets, weights)
            slot[l] = nvec++;
        }
        if (ld < 0) return;
        for (const Ins& I : st.prog)
            if (I.op != I_LEAF && I.op != I_CONST && I.op != I_UNARY && I.op != I_BINARY && I.op != I_POW) return;
        Stage& ds = *plan[st.leaf_stage[ld]];
        if (ds.type != 1 || ds.p != 1 || ds.L.stage >= 0 || ds.R.stage >= 0 || ds.L.adj_mode != 0 || ds.k > 64 || ds.k < 1 ||
            (size_t)ds.m != st.N) return;
        st.fast = 6;
        st.g_dot_leaf = ld; st.g_nvec = nvec; st.g_vec_slot = slot; st.g_dot_stage = st.leaf_stage[ld];
        st.f_rows = ds.m; st.f_cols = ds.k;
    }

    // program text of the one-pass kernel: the stage's constants (mode M_FB) + which leaf is the product, which are row operands
    static bool glm_program_text(Stage& st, std::string& out)
    {
        if (st.enc.empty()) encode_stage(st);
        std::string base;
        if (!jit_program_text(st, M_FB, base)) return false;
        std::ostringstream o;
        // an expensive map is evaluated once per row (rows split over the consumer warps, adjoints through shared memory) instead
        // of once per row and consumer warp: worth its extra barrier from roughly four transcendentals per row on (measured, tools/glm_probe.py)
        int cost = 0;
        for (const DIns& d : st.enc) {
            if (d.code >= F_UNARY && d.code < F_BINARY) {
                const int fn = d.code - F_UNARY;
                cost += (fn == FADB_NEG || fn == FADB_SQUARE || fn == 16) ? 1 : fn == FADB_SQRT ? 12 : 30;
            } else if (d.code == F_BINARY + FADB_DIV || d.code == F_POW || d.code == F_POW2) cost += 12;
            else cost += 1;
        }
        static const char* split_env = getenv("FADB_GLM_SPLIT");
        const bool split = split_env ? split_env[0] == '1' : cost > 100;     // 2^20 x 64: logistic (cost ~65) 0.099 vs 0.100 ms split; four transcendentals + erf 0.200 vs 0.135
        o << base << "#define GLM_SPLIT " << (split ? 1 : 0) << "\n#define GLM_DOT_LEAF " << st.g_dot_leaf << "\n#define GLM_NVEC " << st.g_nvec
          << "\n__device__ constexpr int GLM_VEC_SLOT[JIT_NLEAVES] = {";
        for (size_t l = 0; l < st.g_vec_slot.size(); ++l) o << (l ? ", " : "") << st.g_vec_slot[l];
        o << "};\n";
        out = o.str();
        return true;
    }
    static size_t glm_smem_bytes(const Stage& st)
    {
        const size_t stage_bytes = 64 * 64 * 8 + (size_t)std::max(st.g_nvec, 1) * 64 * 8;
        return 3 * stage_bytes + (64 + 2 * 4 * 64 + 64 + 64 + 64) * sizeof(double) + 2 * 3 * sizeof(unsigned long long);
    }

    // can this call take the one-pass kernel?  (operands may have been rebound since the plan was made)
    bool glm_usable(Context& c, Stage& st)
    {
        if (!g_glm_enabled.load() || !jit_enabled() || st.g_jit_state < 0) return false;
        const Stage& ds = *plan[st.g_dot_stage];
        if (st.f_rows % 2 != 0 || (reinterpret_cast<uintptr_t>(ds.L.val) & 15u)) return false;
        for (size_t l = 0; l < st.leaves.size(); ++l)
            if (st.g_vec_slot[l] >= 0 && (reinterpret_cast<uintptr_t>(st.leaves[l].val) & 15u)) return false;
        if (st.g_jit_state == 1 && st.g_jit.gen != jit_generation()) st.g_jit_state = 0;
        if (st.g_jit_state == 0) {
            std::string text;
            st.g_jit_state = (glm_program_text(st, text) && jit_get(c, text, &st.g_jit, 1, glm_smem_bytes(st)) == FADB_OK) ? 1 : -1;
        }
        return st.g_jit_state == 1;
    }

    int run_glm(Context& c, Stage& st, double seed)
    {
        const Stage& ds = *plan[st.g_dot_stage];
        const int nacc = std::max(st.nch, 1), cols = (int)st.f_cols, nslot = nacc + cols;
        const size_t ntiles = (st.f_rows + 63) / 64;
        const int grid = (int)std::min<size_t>(ntiles, (size_t)c.sms * std::min(std::max(st.g_jit.per_sm, 1), 4));
        const size_t need = (size_t)grid * nslot + nslot;
        if (st.g_ws_cap < need) {
            int rc = dev_alloc(need * sizeof(double), (void**)&st.g_ws);
            if (rc) return rc;
            st.g_ws_cap = need;
        }
        GlmArgsHost a;
        a.rows = st.f_rows; a.cols = cols; a.pad = 0; a.w = ds.R.val; a.seed = seed;
        a.cta_partials = st.g_ws; a.ticket = c.tickets + 20; a.out = st.g_ws + (size_t)grid * nslot;
        for (size_t l = 0; l < (size_t)MAXL_JIT; ++l) a.lp_val[l] = l < st.leaves.size() ? st.leaves[l].val : nullptr;
        alignas(64) unsigned char xmap[128];
        int rc = linreg_encode_xmap(xmap, ds.L.val, st.f_rows, st.f_cols);
        if (rc) return rc;
        void* params[2] = {&a, xmap};
        rc = jit_launch_params(st.g_jit, (unsigned)grid, 160, glm_smem_bytes(st), c.stream, params);
        if (rc) return rc;
        c.launches++;
        GlmFinish f;
        f.nacc = nacc; f.cols = cols; f.nscal = 0; f.term = st.term; f.n_total = (double)st.f_rows;
        f.sig_ptr = st.sig_is_const ? st.d_sig : (st.sig_leaf >= 0 ? st.leaves[st.sig_leaf].val : nullptr);
        f.w_adj = ds.R.adj_mode ? ds.R.adj : nullptr; f.w_mode = ds.R.adj_mode;
        for (const LeafDesc& L : st.leaves)
            if (L.scalar && L.adj_mode != 0 && L.channel > 0 && f.nscal < MAXCH_JIT) {
                f.s_adj[f.nscal] = L.adj; f.s_ch[f.nscal] = L.channel; f.s_mode[f.nscal] = L.adj_mode; ++f.nscal;
            }
        glm_finish_kernel<<<1, 64, 0, c.stream>>>(f, a.out, st.mval);
        c.launches++;
        FADB_CUDA(cudaGetLastError());
        return FADB_OK;
    }

    // fast-path routing of the block plan[first..last] whose root is plan[last] (the whole plan, or one additive term of it)
    void detect_fast_block(size_t first, size_t last)
    {
        const size_t nblk = last - first + 1;
        Stage& st = *plan[last];
        st.fast = 0;
        if (st.type != 0 || nblk > 2) return;
        int lf;
        if (nblk == 2 && st.term != T_NORMAL && st.term != T_SUM) return;
        // sum(pow<2>(y - dot(X, w))) or sum(pow<2>(dot(X, w) - y)), X and y constant: least squares, one pass over X
        if (nblk == 2 && st.term == T_SUM) {
            if (detect_lsq(st)) return;
            detect_glm(st);
            return;
        }
        if (nblk == 1 && st.term == T_SUM && st.prog.size() == 2 && st.root == 1 && leaf_only(st, 0, &lf) &&
            !st.leaves[lf].scalar && (st.prog[1].op == I_UNARY || st.prog[1].op == I_POW) && st.prog[1].a == 0) {
            st.fast = 1;
            st.f_op = st.prog[1].op == I_UNARY ? st.prog[1].fn : 1000 + st.prog[1].fn;
            st.f_x = st.leaves[lf].val; st.f_xadj = st.leaves[lf].adj_mode ? st.leaves[lf].adj : nullptr;
            return;
        }
        if (st.term > T_NORMAL && nblk == 1) {
            int lx, lb, lc = -1;
            const bool three = st.term != T_BERNOULLI;
            if (leaf_only(st, st.nx, &lx) && (!st.leaves[lx].scalar || st.N == 1) && leaf_only(st, st.nm, &lb) &&
                (!three || leaf_only(st, st.ns, &lc)) && st.prog.size() == (three ? 3u : 2u)) {
                auto adj = [&](int l) { return (l >= 0 && st.leaves[l].adj_mode) ? st.leaves[l].adj : nullptr; };
                st.fast = 4;
                st.f_op = st.term - T_CAUCHY;
                st.f_shape = ((!st.leaves[lb].scalar && st.N > 1) ? 1 : 0) | ((three && !st.leaves[lc].scalar && st.N > 1) ? 2 : 0);
                st.f_x = st.leaves[lx].val; st.f_xadj = adj(lx);
                st.f_mu = st.leaves[lb].val; st.f_muadj = adj(lb);
                st.f_sig = three ? st.leaves[lc].val : nullptr; st.f_sigadj = three ? adj(lc) : nullptr;
            }
            return;
        }
        if (st.term == T_NORMAL && st.N >= 1) {
            int lx, lm, ls;
            const bool x_leaf = leaf_only(st, st.nx, &lx) && (!st.leaves[lx].scalar || st.N == 1);
            const bool m_leaf = leaf_only(st, st.nm, &lm);
            const bool s_leaf = leaf_only(st, st.ns, &ls);
            if (nblk == 1 && x_leaf && m_leaf && s_leaf && st.prog.size() == 3) {
                st.fast = 2;
                st.f_shape = (st.mu_vec ? 1 : 0) | (st.sig_vec ? 2 : 0);
                auto adj = [&](int l) { return st.leaves[l].adj_mode ? st.leaves[l].adj : nullptr; };
                st.f_x = st.leaves[lx].val; st.f_xadj = adj(lx);
                st.f_mu = st.leaves[lm].val; st.f_muadj = adj(lm);
                st.f_sig = st.leaves[ls].val; st.f_sigadj = adj(ls);
                return;
            }
            // normal(y const, dot(X const, w var), sigma scalar var): the regression shape
            const Ins& Im = st.prog[st.nm];
            if (x_leaf && s_leaf && !st.sig_vec && st.leaves[lx].adj_mode == 0 && Im.op == I_LEAF &&
                st.leaf_stage[Im.leaf] >= 0 && st.prog.size() == 3) {
                Stage& ds = *plan[st.leaf_stage[Im.leaf]];
                if (ds.type == 1 && ds.p == 1 && ds.L.stage < 0 && ds.R.stage < 0 && ds.L.adj_mode == 0 &&
                    ds.k <= 64 && nblk == 2) {
                    st.fast = 3;
                    st.f_X = ds.L.val; st.f_w = ds.R.val; st.f_wadj = ds.R.adj_mode ? ds.R.adj : nullptr;
                    st.f_rows = ds.m; st.f_cols = ds.k;
                    st.f_x = st.leaves[lx].val;
                    st.f_sig = st.leaves[ls].val; st.f_sigadj = st.leaves[ls].adj_mode ? st.leaves[ls].adj : nullptr;
                }
            }
            // normal(y, f(dot(X, w), ...), sigma) with ONE sigma and a mean that is a map of the product (an intercept, a link
            // function): the same one-pass tile kernel with the log-density as its terminal
            if (st.fast == 0 && nblk == 2 && !st.sig_vec) detect_glm(st);
        }
    }

    // ---- additive roots.  A posterior is a SUM of terms (log-likelihood + log-priors, each a reduction of its own): the
    // adjoint seed of every term is then known before anything is evaluated (the root's sweep only scales the root seed by
    // constants), so each term can run as if it were a root of its own -- forward and adjoint sweep fused in ONE pass, routed to
    // the single-pass kernels where its shape allows -- instead of a forward pass now and a backward pass after the root.
    // Accepted root programs: a scalar stage of + , - , unary minus and * or / by constants over the term values.
    struct TermBlock { int first, last, leaf; };
    std::vector<TermBlock> blocks;          // empty: the plan is executed as one block
    bool analyse_additive_root()
    {
        blocks.clear();
        static const bool off = getenv("FADB_NO_ADDITIVE_ROOT") != nullptr;
        if (off || plan.size() < 2) return false;
        const Stage& rs = *plan.back();
        if (rs.type != 0 || rs.N != 1 || rs.term != T_NONE || rs.big || !rs.assigned.empty()) return false;
        std::vector<double> unused;
        if (!root_term_seeds(rs, 1.0, unused)) return false;
        std::vector<std::pair<int, int>> terms;             // (producer stage, leaf)
        for (size_t l = 0; l < rs.leaves.size(); ++l)
            if (rs.leaf_stage[l] >= 0) terms.push_back({rs.leaf_stage[l], (int)l});
        if (terms.empty()) return false;
        std::sort(terms.begin(), terms.end());
        int prev = -1;
        for (auto& t : terms) {
            if (t.first == prev) return false;              // one producer stage, two leaves: not produced by the planner
            blocks.push_back({prev + 1, t.first, t.second});
            prev = t.first;
        }
        if (prev != (int)plan.size() - 2) { blocks.clear(); return false; }      // every stage below the root belongs to a term
        // ... and to ONE term only: a stage is read by stages of its own block (its term last), a term by the root alone.
        // Anything else (an earlier top-level statement, a placeholder shared between terms) keeps the one-block order.
        const int last = (int)plan.size() - 1;
        std::vector<std::vector<int>> readers(plan.size());
        for (int j = 0; j <= last; ++j) {
            const Stage& st = *plan[j];
            if (!st.assigned.empty() || !st.eq_slot.empty()) { blocks.clear(); return false; }
            if (st.type == 0) { for (int ps : st.leaf_stage) if (ps >= 0) readers[ps].push_back(j); }
            else for (const OperandRef* r : {&st.L, &st.R, &st.S3}) if (r->stage >= 0) readers[r->stage].push_back(j);
        }
        for (const TermBlock& b : blocks)
            for (int k = b.first; k <= b.last; ++k) {
                if (readers[k].empty()) { blocks.clear(); return false; }
                for (int j : readers[k])
                    if (k == b.last ? j != last : (j < b.first || j > b.last)) { blocks.clear(); return false; }
            }
        return true;
    }
    // The root's adjoint sweep on the host: seeds[leaf] = what the root stage would hand to boundary leaf `leaf` for root seed
    // `seed`, with the stage kernel's own operation order.  false: the program is not linear in its term values.
    static bool root_term_seeds(const Stage& rs, double seed, std::vector<double>& seeds)
    {
        const size_t n = rs.prog.size();
        // dep: the slot depends on a term value; isc / cv: the slot is a compile-time constant and its value.  What does not
        // depend on a term (the root's own scalar leaves, whatever is done to them) is the root stage's own business.
        std::vector<char> isc(n, 0), dep(n, 0);
        std::vector<double> cv(n, 0.0), g(n, 0.0);
        for (size_t k = 0; k < n; ++k) {
            const Ins& I = rs.prog[k];
            switch (I.op) {
            case I_CONST: isc[k] = 1; cv[k] = I.imm; break;
            case I_LEAF: dep[k] = rs.leaf_stage[I.leaf] >= 0; break;
            case I_UNARY:
                dep[k] = dep[I.a];
                if (dep[k] && I.fn != FADB_NEG) return false;
                if (I.fn == FADB_NEG) { isc[k] = isc[I.a]; cv[k] = -cv[I.a]; }
                break;
            case I_POW: if (dep[I.a]) return false; break;
            case I_BINARY:
                dep[k] = dep[I.a] || dep[I.b];
                if (dep[k] && I.fn != FADB_ADD && I.fn != FADB_SUB && I.fn != FADB_MUL && I.fn != FADB_DIV) return false;
                if (I.fn == FADB_ADD || I.fn == FADB_SUB || I.fn == FADB_MUL || I.fn == FADB_DIV) {
                    isc[k] = isc[I.a] && isc[I.b];
                    cv[k] = I.fn == FADB_ADD ? cv[I.a] + cv[I.b] : I.fn == FADB_SUB ? cv[I.a] - cv[I.b] : I.fn == FADB_MUL ? cv[I.a] * cv[I.b] : cv[I.a] / cv[I.b];
                }
                break;
            default: return false;                            // placeholders, glue, if_else: one-block order
            }
        }
        if (rs.root < 0 || rs.root >= (int)n) return false;
        seeds.assign(rs.leaves.size(), 0.0);
        g[rs.root] = seed;
        for (int k = (int)n - 1; k >= 0; --k) {
            const Ins& I = rs.prog[k];
            if (!dep[k]) continue;                            // nothing below this slot is a term
            const double sd = g[k];
            switch (I.op) {
            case I_LEAF: seeds[I.leaf] += sd; break;
            case I_UNARY: g[I.a] += -sd; break;
            case I_BINARY:
                if (I.fn == FADB_ADD) { g[I.b] += sd; g[I.a] += sd; }
                else if (I.fn == FADB_SUB) { g[I.b] += -sd; g[I.a] += sd; }
                else if (I.fn == FADB_MUL) {
                    if (dep[I.a] && dep[I.b]) return false;   // a product of two term values: the seeds depend on the values
                    if (dep[I.a]) { if (!isc[I.b]) return false; g[I.a] += sd * cv[I.b]; }      // scaled by a run-time scalar: not known on the host
                    else { if (!isc[I.a]) return false; g[I.b] += sd * cv[I.a]; }
                } else {
                    if (dep[I.b] || !isc[I.b]) return false;  // division by a term value or by a run-time scalar
                    g[I.a] += sd / cv[I.b];
                }
                break;
            default: break;
            }
        }
        return true;
    }

    void detect_fast_paths()
    {
        if (plan.empty()) return;
        for (auto& sp : plan) sp->fast = 0;
        if (analyse_additive_root()) {
            for (const TermBlock& b : blocks) detect_fast_block((size_t)b.first, (size_t)b.last);
            return;
        }
        detect_fast_block(0, plan.size() - 1);
    }

    std::string describe() const
    {
        static const char* opn[] = {"leaf", "const", "unary", "binary", "pow", "eq", "glue", "jz", "jmp", "phi", "opeq"};
        static const char* tn[] = {"store", "sum", "prod", "normal", "cauchy", "uniform", "bernoulli"};
        std::ostringstream o;
        o << "stages=" << plan.size() << " cache={" << cache_vals << "," << cache_adjs << "}\n";
        if (!blocks.empty()) o << "additive root: " << blocks.size() << " term(s), each swept as a root of its own\n";
        for (size_t k = 0; k < plan.size(); ++k) {
            const Stage& st = *plan[k];
            if (st.type == 5) { o << "stage " << k << ": wishart p=" << st.m << " df=" << st.sig_const << "\n"; continue; }
            if (st.type == 4) { o << "stage " << k << ": " << (st.k ? "log_det" : "det") << " n=" << st.m << " policy=" << st.p << "\n"; continue; }
            if (st.type == 3) { o << "stage " << k << ": dense normal n=" << st.m << (st.k ? " (vvm)" : " (vsm)") << "\n"; continue; }
            if (st.type == 2) { o << "stage " << k << ": transpose " << st.m << "x" << st.k << "\n"; continue; }
            if (st.type == 1) { o << "stage " << k << ": dot " << st.m << "x" << st.k << " * " << st.k << "x" << st.p << "\n"; continue; }
            o << "stage " << k << ": map N=" << st.N << " term=" << tn[st.term] << " ins=" << st.prog.size()
              << " leaves=" << st.leaves.size() << " channels=" << st.nch << " fast=" << st.fast
              << (st.big ? " big" : "") << (st.jit_only ? " specialised-only" : "") << "\n";
            for (size_t i = 0; i < st.prog.size() && i < 256; ++i) {
                const Ins& I = st.prog[i];
                o << "  %" << i << " = " << opn[I.op];
                if (I.op == I_LEAF) o << " L" << I.leaf << (st.leaves[I.leaf].scalar ? " (broadcast)" : "")
                                      << (st.leaf_stage[I.leaf] >= 0 ? " <- stage " + std::to_string(st.leaf_stage[I.leaf]) : "");
                else if (I.op == I_CONST) o << " " << I.imm;
                else if (I.op == I_UNARY) o << "[" << I.fn << "] %" << I.a;
                else if (I.op == I_POW) o << "<" << I.fn << "> %" << I.a;
                else if (I.op == I_BINARY) o << "[" << I.fn << "] %" << I.a << " %" << I.b;
                else if (I.op == I_EQ) o << " L" << I.leaf << " := %" << I.a;
                else if (I.op == I_OPEQ) o << "[" << I.fn << "] L" << I.leaf << " (old %" << I.b << ") op= %" << I.a;
                else if (I.op == I_JZ) o << " %" << I.a << " -> @" << I.b;
                else if (I.op == I_JMP) o << " -> @" << I.b;
                else if (I.op == I_PHI) o << " %" << I.leaf << " ? %" << I.a << " : %" << I.b;
                else o << " %" << I.a << " , %" << I.b;
                o << "\n";
            }
        }
        return o.str();
    }

    // device encoding of a stage program (host only; needs the planner's leaf table)
    static void encode_stage(Stage& st)
    {
        std::vector<DIns>& enc = st.enc;
        enc.assign(st.prog.size(), DIns{});
        for (size_t k = 0; k < st.prog.size(); ++k) {
            const Ins& I = st.prog[k];
            DIns D{0, (short)I.a, (short)I.b, (short)I.leaf, I.imm};
            switch (I.op) {
            case I_LEAF: D.code = st.leaves[I.leaf].scalar ? F_LEAF_S : F_LEAF_V; break;
            case I_CONST: D.code = F_CONST; break;
            case I_EQ: D.code = F_EQ; break;
            case I_GLUE: D.code = F_GLUE; break;
            case I_POW: D.code = I.fn == 2 ? F_POW2 : F_POW; D.leaf = (short)I.fn; break;
            case I_UNARY: D.code = (short)(F_UNARY + (I.fn == FADB_MOCK2X ? 16 : I.fn)); break;
            case I_BINARY: D.code = (short)(F_BINARY + (I.fn == FADB_MOCKB ? 12 : I.fn)); break;
            case I_JZ: D.code = F_JZ; break;
            case I_JMP: D.code = F_JMP; break;
            case I_PHI: D.code = F_PHI; break;
            default: D.code = (short)(F_OPEQ + I.fn); break;
            }
            enc[k] = D;
        }
    }

    // ---------------------------------------------------------------- device state
    int dev_alloc(size_t bytes, void** out)
    {
        FADB_CUDA(cudaMalloc(out, bytes ? bytes : 8));
        owned.push_back(*out);
        return FADB_OK;
    }

    int materialize(Context& c)
    {
        for (size_t k = 0; k < plan.size(); ++k) {
            Stage& st = *plan[k];
            int rc;
            if ((rc = dev_alloc(st.out_size * sizeof(double), (void**)&st.mval))) return rc;
            if ((rc = dev_alloc(st.out_size * sizeof(double), (void**)&st.madj))) return rc;
            FADB_CUDA(cudaMemsetAsync(st.madj, 0, st.out_size * sizeof(double), c.stream));
            if ((rc = dev_alloc(8 * sizeof(double), (void**)&st.d_red))) return rc;
            FADB_CUDA(cudaMemsetAsync(st.d_red, 0, 8 * sizeof(double), c.stream));
        }
        for (auto& sp : plan) {
            Stage& st = *sp;
            if (st.type != 0) {
                if (st.type == 1 && st.p <= 16 && st.m >= 4096 && st.k <= 4096) {
                    // row blocks of >= 1024 rows (4 per thread), as many as keep ~4 CTAs per SM busy with the column tiles
                    {
                        const int ctiles = (st.k + GV_KT - 1) / GV_KT;
                        const int want = std::max(1, 4 * c.sms / ctiles);
                        st.gv_grid = std::max(1, std::min(want, (st.m + 1023) / 1024));
                    }
                    int rc = dev_alloc((size_t)st.gv_grid * st.k * sizeof(double), (void**)&st.gv_part);
                    if (rc) return rc;
                } else if (st.type == 1 && st.p >= 4096 && st.m * st.k <= GV_THREADS * WR_PER && st.L.adj_mode) {
                    st.wr_grid = (int)std::min<size_t>(2 * c.sms, ((size_t)st.p + 1023) / 1024);
                    int rc = dev_alloc((size_t)st.wr_grid * st.m * st.k * sizeof(double), (void**)&st.wr_part);
                    if (rc) return rc;
                }
                if (st.L.stage >= 0) { st.L.val = plan[st.L.stage]->mval; st.L.adj = plan[st.L.stage]->madj; }
                if ((st.type == 1 || st.type == 3 || st.type == 5) && st.R.stage >= 0) { st.R.val = plan[st.R.stage]->mval; st.R.adj = plan[st.R.stage]->madj; }
                if (st.type == 3 && st.S3.stage >= 0) { st.S3.val = plan[st.S3.stage]->mval; st.S3.adj = plan[st.S3.stage]->madj; }
                continue;
            }
            for (size_t l = 0; l < st.leaves.size(); ++l)
                if (st.leaf_stage[l] >= 0) {
                    st.leaves[l].val = plan[st.leaf_stage[l]]->mval;
                    st.leaves[l].adj = plan[st.leaf_stage[l]]->madj;
                }
            int rc;
            if ((rc = dev_alloc(std::max<size_t>(1, st.prog.size()) * sizeof(DIns), (void**)&st.d_prog))) return rc;
            if ((rc = dev_alloc(std::max<size_t>(1, st.leaves.size()) * sizeof(LeafDesc), (void**)&st.d_leaves))) return rc;
            encode_stage(st);
            std::vector<DIns>& enc = st.enc;
            FADB_CUDA(cudaMemcpyAsync(st.d_prog, enc.data(), enc.size() * sizeof(DIns), cudaMemcpyHostToDevice, c.stream));
            FADB_CUDA(cudaMemcpyAsync(st.d_leaves, st.leaves.data(), st.leaves.size() * sizeof(LeafDesc), cudaMemcpyHostToDevice, c.stream));
            // pre-stage values of leaves that are read and then assigned in this stage (see Stage::pre)
            st.pre.clear();
            for (size_t l = 0; l < st.leaves.size(); ++l) {
                if (!st.leaves[l].assigned || st.leaf_node[l] < 0) continue;
                bool read = false, opeq = false;
                for (const Ins& I : st.prog) {
                    if (I.op == I_LEAF && I.leaf == (int)l) read = true;
                    if (I.op == I_OPEQ && I.leaf == (int)l) opeq = true;
                }
                if (!read) continue;
                Stage::PreVal pv{(int)l, nullptr, st.leaves[l].scalar ? (size_t)1 : st.N, opeq};
                if ((rc = dev_alloc(pv.n * sizeof(double), (void**)&pv.cache))) return rc;
                st.pre.push_back(pv);
            }
            if (!st.pre.empty()) {
                if ((rc = dev_alloc(st.leaves.size() * sizeof(LeafDesc), (void**)&st.d_leaves_b))) return rc;
                if ((rc = upload_leaves_b(c, st))) return rc;
            }
            if (st.term == T_NORMAL && st.sig_is_const) {
                if ((rc = dev_alloc(sizeof(double), (void**)&st.d_sig))) return rc;
                FADB_CUDA(cudaMemcpyAsync(st.d_sig, &st.sig_const, sizeof(double), cudaMemcpyHostToDevice, c.stream));
            }
            if (st.fast == 3 || st.fast == 5) {
                if ((rc = dev_alloc((st.f_cols + 1) * sizeof(double), (void**)&st.f_part))) return rc;
            }
            if (st.big) {
                if ((rc = dev_alloc((2 * st.prog.size() + st.leaves.size()) * sizeof(double), (void**)&st.d_scratch))) return rc;
            }
        }
        FADB_CUDA(cudaStreamSynchronize(c.stream));   // host vectors may be reallocated later
        bound = true;
        return FADB_OK;
    }

    // d_leaves_b = leaf table of the backward-only launches; the caches start as a copy of the current values
    int upload_leaves_b(Context& c, Stage& st)
    {
        std::vector<LeafDesc> lb = st.leaves;
        for (const Stage::PreVal& pv : st.pre) {
            lb[pv.leaf].val = pv.cache;
            FADB_CUDA(cudaMemcpyAsync(pv.cache, st.leaves[pv.leaf].val, pv.n * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
        }
        FADB_CUDA(cudaMemcpyAsync(st.d_leaves_b, lb.data(), lb.size() * sizeof(LeafDesc), cudaMemcpyHostToDevice, c.stream));
        FADB_CUDA(cudaStreamSynchronize(c.stream));      // `lb` is a local
        return FADB_OK;
    }

    int reupload_leaves(Context& c)
    {
        for (auto& sp : plan) {
            Stage& st = *sp;
            if (st.type != 0) {
                if (st.L.stage < 0) { st.L.val = nodes[st.L.node].val; st.L.adj = nodes[st.L.node].adj; }
                if ((st.type == 1 || st.type == 3 || st.type == 5) && st.R.stage < 0) { st.R.val = nodes[st.R.node].val; st.R.adj = nodes[st.R.node].adj; }
                if (st.type == 3 && st.S3.stage < 0) { st.S3.val = nodes[st.S3.node].val; st.S3.adj = nodes[st.S3.node].adj; st.factored = false; }
                continue;
            }
            for (size_t l = 0; l < st.leaves.size(); ++l)
                if (st.leaf_node[l] >= 0) {
                    st.leaves[l].val = nodes[st.leaf_node[l]].val;
                    st.leaves[l].adj = st.leaves[l].adj_mode ? nodes[st.leaf_node[l]].adj : nullptr;
                }
            FADB_CUDA(cudaMemcpyAsync(st.d_leaves, st.leaves.data(), st.leaves.size() * sizeof(LeafDesc), cudaMemcpyHostToDevice, c.stream));
            if (!st.pre.empty()) { int rc = upload_leaves_b(c, st); if (rc) return rc; }
        }
        FADB_CUDA(cudaStreamSynchronize(c.stream));
        detect_fast_paths();
        return FADB_OK;
    }

    // ---------------------------------------------------------------- specialisation (jit.cu)
    // The compile-time half of a stage: program, leaf attributes, terminal configuration, mode.
    // Pointers, N and seeds are NOT part of it.  false: this stage stays on the interpreter.
    // 4 elements per thread with 256-bit accesses: per-element operands only through F_LEAF_V (no
    // placeholder or op= stores, whose order inside the program matters), and enough elements
    static int jit_vec_width(const Stage& st)
    {
        static const bool off = getenv("FADB_JIT_NO_VEC") != nullptr;
        if (off || st.N < 4096) return 1;
        bool any_vec = false;
        for (const LeafDesc& l : st.leaves) {
            if (l.assigned) return 1;
            any_vec = any_vec || !l.scalar;
        }
        for (const DIns& d : st.enc)
            if (d.code == F_EQ || d.code >= F_OPEQ) return 1;
        return any_vec ? 4 : 1;
    }
    // The operand ring (stage_kernel.cuh, JIT_RING) serves the one-element form of per-element stages with placeholder
    // stores: no op= (its stores feed later reads of the same leaf through memory) and no if_else.
    static bool jit_ring_eligible(const Stage& st)
    {
        static const bool off = getenv("FADB_JIT_NO_RING") != nullptr;
        if (off || !g_ring_enabled.load() || st.big || st.N < 4096 || !st.pre.empty() || jit_vec_width(st) != 1) return false;
        bool any_vec = false;
        for (const LeafDesc& l : st.leaves) any_vec = any_vec || !l.scalar;
        for (const DIns& d : st.enc)
            if (d.code >= F_OPEQ || d.code == F_JZ || d.code == F_JMP || d.code == F_PHI) return false;
        return any_vec;
    }
    // A copy issued one trip early must not be overtaken by a store: every buffer the stage WRITES (placeholder values,
    // adjoints, its own output) has to be disjoint from every other per-element buffer it touches.
    static bool ring_buffers_disjoint(const Stage& st, const StageArgs& a)
    {
        struct Span { const char* lo; const char* hi; bool w; };
        std::vector<Span> sp;
        const size_t bytes = st.N * sizeof(double);
        auto add = [&](const void* p, bool w) { if (p) sp.push_back({(const char*)p, (const char*)p + bytes, w}); };
        for (size_t l = 0; l < st.leaves.size(); ++l) {
            const LeafDesc& d = st.leaves[l];
            if (d.scalar) continue;
            add(a.lp_val[l], d.assigned != 0);
            if (d.adj_mode) add(a.lp_adj[l], true);
        }
        add(a.out_vec, true);
        add(a.seed_vec, false);
        for (size_t i = 0; i < sp.size(); ++i)
            for (size_t j = i + 1; j < sp.size(); ++j)
                if ((sp[i].w || sp[j].w) && sp[i].lo < sp[j].hi && sp[j].lo < sp[i].hi) return false;
        return true;
    }
    static bool jit_program_text(const Stage& st, int mode, std::string& out, bool ring = false)
    {
        if (st.big || st.enc.empty() || st.enc.size() > (size_t)MAXS_JIT || st.leaves.size() > (size_t)MAXL_JIT) return false;
        if (ring && !jit_ring_eligible(st)) return false;
        std::ostringstream o;
        char buf[64];
        // registers: 8 CTAs/SM (64 registers) is the measured optimum for small programs; larger ones spill there
        if (st.enc.size() > 64) o << "\n#ifndef FADB_JIT_MINB\n#define FADB_JIT_MINB " << (st.enc.size() > 128 ? 2 : 4) << "\n#endif";
        bool cf = false;
        for (const DIns& d : st.enc) cf = cf || d.code == F_JZ || d.code == F_JMP || d.code == F_PHI;
        // if_else: the unrolled loops carry a jump watermark (stage_kernel.cuh); values stay live across the guarded regions, so
        // the 64-register cap of small programs would spill (544 bytes on a 25-node nested case): 128 registers instead
        if (cf) o << "\n#define JIT_HAS_CF 1" << (st.enc.size() <= 64 ? "\n#ifndef FADB_JIT_MINB\n#define FADB_JIT_MINB 4\n#endif" : "");
        o << "\n#define JIT_VEC " << jit_vec_width(st) << "\n#define JIT_NINS " << st.enc.size() << "\n#define JIT_NLEAVES " << st.leaves.size()
          << "\n#define JIT_MODE " << mode << "\n#define JIT_TERM " << st.term << "\n#define JIT_ROOT " << st.root
          << "\n#define JIT_NX " << st.nx << "\n#define JIT_NM " << st.nm << "\n#define JIT_NS " << st.ns
          << "\n#define JIT_MU_VEC " << st.mu_vec << "\n#define JIT_SIG_VEC " << st.sig_vec << "\n#define JIT_NCH " << st.nch
          << "\n__device__ constexpr DIns JIT_PROG[JIT_NINS] = {\n";
        // ring form: a per-element read behind the leaf's assignment names the assigning instruction in `a` (-1: from memory)
        const size_t nl = st.leaves.size();
        std::vector<int> last_eq(nl, -1), rv(nl, -1), ra(nl, -1);
        std::vector<char> reads_mem(nl, 0);
        for (size_t k = 0; k < st.enc.size(); ++k) {
            const DIns& d = st.enc[k];
            if (!std::isfinite(d.imm)) return false;
            snprintf(buf, sizeof buf, "%a", d.imm);                                      // hex float: exact
            int fa = d.a;
            if (ring && d.code == F_LEAF_V) {
                fa = last_eq[d.leaf];
                if (fa < 0) reads_mem[d.leaf] = 1;
            }
            if (ring && d.code == F_EQ && !st.leaves[d.leaf].scalar) last_eq[d.leaf] = (int)k;
            o << "  {" << d.code << ", " << fa << ", " << d.b << ", " << d.leaf << ", " << buf << "},\n";
        }
        if (ring) {
            int nrv = 0, nra = 0;
            for (size_t l = 0; l < nl; ++l) {
                if (st.leaves[l].scalar) continue;
                if (reads_mem[l]) rv[l] = nrv++;
                if (st.leaves[l].adj_mode == 1 && (mode & M_B)) ra[l] = nra++;
            }
            if (nrv + nra == 0 || nrv + nra + 1 > 40) return false;                      // 1 KB of shared memory per slot and CTA
            o << "};\n#define JIT_RING 1\n#define JIT_NRV " << nrv << "\n#define JIT_NRA " << nra << "\n#define JIT_NRS " << (st.term == T_NONE ? 1 : 0)
              << "\n__device__ constexpr int JIT_RV[] = {";
            for (size_t l = 0; l < nl; ++l) o << rv[l] << ", ";
            o << "};\n__device__ constexpr int JIT_RA[] = {";
            for (size_t l = 0; l < nl; ++l) o << ra[l] << ", ";
            o << "};\n__device__ constexpr JitLeaf JIT_LEAF[JIT_NLEAVES > 0 ? JIT_NLEAVES : 1] = {\n";
        } else
            o << "};\n__device__ constexpr JitLeaf JIT_LEAF[JIT_NLEAVES > 0 ? JIT_NLEAVES : 1] = {\n";
        for (const LeafDesc& l : st.leaves)
            o << "  {" << l.scalar << ", " << l.adj_mode << ", " << l.channel << ", " << l.assigned << "},\n";
        if (st.leaves.empty()) o << "  {0, 0, -1, 0}\n";
        o << "};\n";
        out = o.str();
        return true;
    }

    // ---------------------------------------------------------------- execution
    int run_map(Context& c, Stage& st, int mode, double seed, const double* seed_vec, const double* seed_ptr)
    {
        StageArgs a;
        a.prog = st.d_prog; a.leaves = st.d_leaves;
        a.nins = (int)st.prog.size(); a.nleaves = (int)st.leaves.size();
        a.N = st.N; a.mode = mode; a.term = st.term; a.root = st.root;
        a.nx = st.nx; a.nm = st.nm; a.ns = st.ns; a.mu_vec = st.mu_vec; a.sig_vec = st.sig_vec;
        a.sig_ptr = st.sig_is_const ? st.d_sig : (st.sig_leaf >= 0 ? st.leaves[st.sig_leaf].val : nullptr);
        a.out_vec = st.term == T_NONE ? st.mval : nullptr;
        a.out_scalar = st.mval;
        a.seed_vec = seed_vec; a.seed_ptr = seed_ptr; a.seed = seed;
        a.red = st.d_red;
        a.gate = (((st.term == T_NORMAL && st.sig_vec) || (st.term > T_NORMAL && st.needs_check)) && mode == M_B) ? st.d_red + 1 : nullptr;
        a.nch = st.nch; a.term_ch = st.term_ch;
        a.partials = c.partials; a.ticket = c.tickets + 8;
        a.scratch = st.d_scratch;
        a.vec_ok = 0;
        for (size_t l = 0; l < st.leaves.size() && l < (size_t)MAXL_JIT; ++l) { a.lp_val[l] = st.leaves[l].val; a.lp_adj[l] = st.leaves[l].adj; }
        if (!st.pre.empty() && mode == M_F)        // forward-only launch: keep the values it is about to overwrite
            for (const Stage::PreVal& pv : st.pre)
                FADB_CUDA(cudaMemcpyAsync(pv.cache, st.leaves[pv.leaf].val, pv.n * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
        if (!st.pre.empty() && mode == M_B) {      // backward-only launch: the recomputed forward pass reads them back
            a.leaves = st.d_leaves_b;
            for (const Stage::PreVal& pv : st.pre) a.lp_val[pv.leaf] = pv.cache;
        }
        int rc_launch = launch_map(c, st, mode, a);
        if (rc_launch) return rc_launch;
        if (!st.pre.empty() && mode == M_B)        // OpEqNode::beval leaves the previous value in w (eq.hpp:262-271)
            for (const Stage::PreVal& pv : st.pre)
                if (pv.restore)
                    FADB_CUDA(cudaMemcpyAsync(const_cast<double*>(st.leaves[pv.leaf].val), pv.cache, pv.n * sizeof(double),
                                              cudaMemcpyDeviceToDevice, c.stream));
        return FADB_OK;
    }

    int launch_map(Context& c, Stage& st, int mode, StageArgs& a)
    {
        if (st.big) {
            stage_kernel<2><<<1, 32, 0, c.stream>>>(a);
            c.launches++;
            FADB_CUDA(cudaGetLastError());
            return FADB_OK;
        }
        if (jit_enabled() && st.jit_ring_state[mode] >= 0 && jit_ring_eligible(st) && ring_buffers_disjoint(st, a)) {
            // the operand-ring form of the specialised kernel (placeholder stages), when no written operand overlaps another
            if (st.jit_ring_state[mode] == 1 && st.jit_ring[mode].gen != jit_generation()) st.jit_ring_state[mode] = 0;
            if (st.jit_ring_state[mode] == 0) {
                std::string text;
                st.jit_ring_state[mode] = (jit_program_text(st, mode, text, true) && jit_get(c, text, &st.jit_ring[mode]) == FADB_OK) ? 1 : -1;
            }
            if (st.jit_ring_state[mode] == 1) {
                a.vec_ok = 1;
                int grid = grid_for(c, st.N, st.jit_ring[mode].per_sm, 128);
                const int kacc = std::max(st.nch, 3);
                if ((size_t)grid * (kacc + 1) > (size_t)kMaxGrid * kPartialSlots) grid = kMaxGrid * kPartialSlots / (kacc + 1);
                int rc = jit_launch(st.jit_ring[mode], (unsigned)grid, 128, c.stream, &a);
                if (rc) return rc;
                c.launches++;
                return FADB_OK;
            }
        }
        if (jit_enabled() && st.jit_state[mode] >= 0) {
            // specialised kernel of this stage program (jit.cu): compiled once per (program, mode, device)
            if (st.jit_state[mode] == 1 && st.jit[mode].gen != jit_generation()) st.jit_state[mode] = 0;   // unloaded by fadb_shutdown
            if (st.jit_state[mode] == 0) {
                std::string text;
                st.jit_state[mode] = (jit_program_text(st, mode, text) && jit_get(c, text, &st.jit[mode]) == FADB_OK) ? 1 : -1;
            }
            if (st.jit_state[mode] == 1) {
                const int vw = jit_vec_width(st);
                a.vec_ok = 1;
                if (vw > 1)
                    for (const LeafDesc& l : st.leaves)
                        if (!l.scalar && (!aligned32(l.val) || (l.adj_mode && !aligned32(l.adj)))) a.vec_ok = 0;
                int grid = grid_for(c, (st.N + vw - 1) / vw, st.jit[mode].per_sm, 128);
                const int kacc = std::max(st.nch, 3);      // K_ACC of the specialised kernel
                if ((size_t)grid * (kacc + 1) > (size_t)kMaxGrid * kPartialSlots) grid = kMaxGrid * kPartialSlots / (kacc + 1);
                int rc = jit_launch(st.jit[mode], (unsigned)grid, 128, c.stream, &a);
                if (rc) return rc;
                c.launches++;
                return FADB_OK;
            }
        }
        if (st.jit_only) {
            set_error("fastad_b200: this expression stage (" + std::to_string(st.prog.size()) + " nodes, " + std::to_string(st.leaves.size()) +
                      " leaves, " + std::to_string(st.nch) + " reduction channels) exceeds the interpreter's limits and needs the "
                      "NVRTC-specialised kernel: " + fadb_jit_diagnostics());
            return FADB_ERR_UNSUPPORTED;
        }
        const int threads = st.N >= 4096 ? 128 : 32;
        const size_t smem = (2 * st.prog.size() + st.leaves.size()) * threads * sizeof(double);
        // Measured on B200 (profiles/): shared-memory slots are NOT a win -- 0.99 vs 1.05 ms on a
        // 10-node chain, and 3x slower on the 50-node Black-Scholes stage where 111 KB per CTA
        // collapses occupancy; the interpreter is issue-bound, not local-memory-bound.  Local
        // arrays stay the default; FADB_STAGE_SMEM=1 selects the shared-memory layout.
        static const bool use_local = getenv("FADB_STAGE_SMEM") == nullptr;
        if (use_local) {
            int grid = resident_grid(c, stage_kernel<0>, threads, 0, st.N, threads);
            if ((size_t)grid * (MAXCH + 1) > (size_t)kMaxGrid * kPartialSlots) grid = kMaxGrid * kPartialSlots / (MAXCH + 1);
            stage_kernel<0><<<grid, threads, 0, c.stream>>>(a);
        } else {
            static bool attr_set[kMaxDevices] = {false};
            if (!attr_set[c.device]) {
                FADB_CUDA(cudaFuncSetAttribute(stage_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)((2 * MAXS + MAXL) * 128 * sizeof(double))));
                attr_set[c.device] = true;
            }
            int grid = resident_grid(c, stage_kernel<1>, threads, smem, st.N, threads);
            if ((size_t)grid * (MAXCH + 1) > (size_t)kMaxGrid * kPartialSlots) grid = kMaxGrid * kPartialSlots / (MAXCH + 1);
            stage_kernel<1><<<grid, threads, smem, c.stream>>>(a);
        }
        c.launches++;
        FADB_CUDA(cudaGetLastError());
        return FADB_OK;
    }

    int run_dot(Context& c, Stage& st, int mode, const double* seed_vec, double seed)
    {
        const int threads = 128;
        if (st.type >= 3) {
            // forward-only evaluations zero the seed (no adjoint is touched, normal.hpp:644); an inner
            // stage reads its scalar seed back from its boundary buffer.  The dense workspace is shared by
            // all dense stages of a device: a backward-only call re-factors unless this stage used it last.
            static unsigned long long owner[kMaxDevices] = {};      // Stage::uid (never reused, unlike a freed stage's address)
            double sd = (mode & M_B) ? seed : 0.0;
            if (mode == M_B && !seed_vec) {
                FADB_CUDA(cudaMemcpyAsync(&sd, st.madj, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
                FADB_CUDA(cudaStreamSynchronize(c.stream));
            }
            const bool mine = owner[c.device] == st.uid;
            auto adjp = [](const OperandRef& r) { return r.adj_mode ? r.adj : nullptr; };
            int rc;
            if (st.type == 3) {
                const bool refactor = !(mine && ((st.sigma_const && st.factored) || mode == M_B));
                rc = fadb_normal_dense_async(st.m, st.L.val, st.R.val, st.k, st.S3.val, adjp(st.L), adjp(st.R), adjp(st.S3), sd,
                                             0, refactor ? 1 : 0, st.mval);
                st.factored = true;
            } else if (st.type == 4) {
                rc = fadb_det_async(st.m, st.L.val, adjp(st.L), sd, 0, st.p, st.k, (mine && mode == M_B) ? 0 : 1, st.mval);
            } else {
                rc = fadb_wishart_async(st.m, st.L.val, st.R.val, st.sig_const, adjp(st.L), adjp(st.R), sd, 0,
                                        (mine && mode == M_B) ? 0 : 1, st.mval);
            }
            if (rc) return rc;
            owner[c.device] = st.uid;
            return FADB_OK;
        }
        if (st.type == 2) {
            const int n = st.m * st.k, blocks = (n + threads - 1) / threads;
            if (n == 0) return FADB_OK;
            if (mode & M_F) {
                transpose_fwd_kernel<<<blocks, threads, 0, c.stream>>>(st.m, st.k, st.L.val, st.mval);
                c.launches++;
            }
            if ((mode & M_B) && st.L.adj_mode) {
                if (!seed_vec && mode == M_FB) {
                    fill_kernel<<<blocks, threads, 0, c.stream>>>(st.madj, n, seed);
                    c.launches++;
                }
                transpose_bwd_kernel<<<blocks, threads, 0, c.stream>>>(st.m, st.k, seed_vec ? seed_vec : st.madj, st.L.adj,
                                                                        st.L.adj_mode);
                c.launches++;
            }
            FADB_CUDA(cudaGetLastError());
            return FADB_OK;
        }
        if (st.gv_part) {       // tall L, R a vector or a few columns: the HBM-bound kernels, one launch per column of R
            const size_t mp = (size_t)st.m * st.p;
            if (mode & M_F) {
                const bool split = (size_t)st.m < (size_t)c.sms * 8 * GV_THREADS;      // too few rows for one thread each
                for (int j = 0; j < st.p; ++j) {
                    if (split)
                        gemv_n_kernel<4><<<grid_for(c, st.m, 8, GV_THREADS / 4), GV_THREADS, (st.k + GV_THREADS) * sizeof(double), c.stream>>>(
                            (size_t)st.m, st.k, st.L.val, st.R.val + (size_t)j * st.k, st.mval + (size_t)j * st.m);
                    else
                        gemv_n_kernel<1><<<grid_for(c, st.m, 8, GV_THREADS), GV_THREADS, st.k * sizeof(double), c.stream>>>(
                            (size_t)st.m, st.k, st.L.val, st.R.val + (size_t)j * st.k, st.mval + (size_t)j * st.m);
                    c.launches++;
                }
            }
            if (mode & M_B) {
                const double* A = seed_vec ? seed_vec : st.madj;
                if (!seed_vec && mode == M_FB) {
                    fill_kernel<<<(unsigned)((mp + threads - 1) / threads), threads, 0, c.stream>>>(st.madj, mp, seed);
                    c.launches++;
                }
                if (st.R.adj_mode)
                    for (int j = 0; j < st.p; ++j) {
                        gemv_t_kernel<<<dim3(st.gv_grid, (st.k + GV_KT - 1) / GV_KT), GV_THREADS, 0, c.stream>>>((size_t)st.m, st.k, st.L.val, A + (size_t)j * st.m, st.gv_part,
                                                                                c.tickets + 12, st.R.adj + (size_t)j * st.k, st.R.adj_mode);
                        c.launches++;
                    }
                if (st.L.adj_mode) {
                    if (st.p == 1)
                        outer_kernel<<<grid_for(c, (size_t)st.m * st.k, 8, 256), 256, 0, c.stream>>>((size_t)st.m, st.k, A, st.R.val, st.L.adj,
                                                                                                      st.L.adj_mode);
                    else {
                        const int n = st.m * st.k;
                        dot_bwd_kernel<<<(n + threads - 1) / threads, threads, 0, c.stream>>>(st.m, st.k, st.p, st.L.val, st.R.val, A, st.L.adj,
                                                                                                st.L.adj_mode, nullptr, 0);
                    }
                    c.launches++;
                }
            }
            FADB_CUDA(cudaGetLastError());
            return FADB_OK;
        }
        // matrix x matrix with every dimension large enough for 64-wide tiles: the double-buffered FP64 GEMM
        const bool tiled = st.m >= 64 && st.p >= 64 && st.k >= 32;
        if (tiled) {
            if (mode & M_F) gemm_f64(c, false, false, st.m, st.p, st.k, 1.0, st.L.val, st.m, st.R.val, st.k, 0.0, st.mval, st.m);   // V = L R
            if (mode & M_B) {
                const double* A = seed_vec ? seed_vec : st.madj;
                if (!seed_vec && mode == M_FB) {
                    const size_t n = (size_t)st.m * st.p;
                    fill_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, c.stream>>>(st.madj, n, seed);
                    c.launches++;
                }
                if (st.R.adj_mode)      // R_adj (+)= L^T A   (dot.hpp:100)
                    gemm_f64(c, true, false, st.k, st.p, st.m, 1.0, st.L.val, st.m, A, st.m, st.R.adj_mode == 2 ? 0.0 : 1.0, st.R.adj, st.k);
                if (st.L.adj_mode)      // L_adj (+)= A R^T   (dot.hpp:101)
                    gemm_f64(c, false, true, st.m, st.k, st.p, 1.0, A, st.m, st.R.val, st.k, st.L.adj_mode == 2 ? 0.0 : 1.0, st.L.adj, st.m);
            }
            FADB_CUDA(cudaGetLastError());
            return FADB_OK;
        }
        if (mode & M_F) {
            int n = st.m * st.p;
            dot_fwd_kernel<<<(n + threads - 1) / threads, threads, 0, c.stream>>>(st.m, st.k, st.p, st.L.val, st.R.val, st.mval);
            c.launches++;
        }
        if (mode & M_B) {
            const double* A = seed_vec ? seed_vec : st.madj;
            if (!seed_vec && mode == M_FB) {      // root dot with a scalar seed: broadcast it
                int n = st.m * st.p;
                fill_kernel<<<(n + threads - 1) / threads, threads, 0, c.stream>>>(st.madj, n, seed);
                c.launches++;
            }
            const bool wide = st.wr_part != nullptr;    // A R^T summed over a large p: partial sums per CTA
            if (wide) {
                if (st.m * st.k <= WS_MAXN)
                    wide_outer_small_kernel<<<st.wr_grid, GV_THREADS, 0, c.stream>>>(st.m, st.k, (size_t)st.p, A, st.R.val, st.wr_part,
                                                                                      c.tickets + 13, st.L.adj, st.L.adj_mode);
                else
                    wide_outer_kernel<<<st.wr_grid, GV_THREADS, 0, c.stream>>>(st.m, st.k, (size_t)st.p, A, st.R.val, st.wr_part, c.tickets + 13,
                                                                                st.L.adj, st.L.adj_mode);
                c.launches++;
            }
            const bool want_l = st.L.adj_mode && !wide;
            if (want_l || st.R.adj_mode) {
                int n = std::max(want_l ? st.m * st.k : 0, st.R.adj_mode ? st.k * st.p : 0);
                dot_bwd_kernel<<<(n + threads - 1) / threads, threads, 0, c.stream>>>(
                    st.m, st.k, st.p, st.L.val, st.R.val, A, want_l ? st.L.adj : nullptr, st.L.adj_mode,
                    st.R.adj_mode ? st.R.adj : nullptr, st.R.adj_mode);
                c.launches++;
            }
        }
        FADB_CUDA(cudaGetLastError());
        return FADB_OK;
    }

    int run_fast_root(Context& c, Stage& st, double seed)
    {
        if (st.fast == 1)
            return fadb_sum_unary_async(st.f_op, st.N, st.f_x, st.f_xadj, seed, 0, st.mval);
        if (st.fast == 2) {
            // vector sigma: optimistic pass + undo behind the finish (normal.cu), or the pre-pass with FADB_STRICT_SIGMA
            const bool guarded = st.sig_vec && seed != 0.0 && (st.f_xadj || st.f_muadj || st.f_sigadj) && !(flags & FADB_TRUST_SIGMA);
            const double* inv = nullptr;
            if (guarded && (flags & FADB_STRICT_SIGMA)) {
                int rc = fadb_sigma_check(st.N, st.f_sig, st.d_red + 1);
                if (rc) return rc;
                inv = st.d_red + 1;
            }
            const bool sss = st.N == 1 && st.f_shape == 0;
            int rc = fadb_normal_lpdf_partial(st.f_shape, st.N, st.f_x, st.f_mu, st.f_sig, sss ? nullptr : st.f_xadj,
                                              sss ? nullptr : st.f_muadj, sss ? nullptr : st.f_sigadj, seed, 0, inv,
                                              st.d_red + 4);
            if (rc) return rc;
            rc = fadb_normal_lpdf_finish(st.f_shape, st.N, st.d_red + 4, st.f_x, st.f_mu, st.f_sig,
                                         sss ? st.f_xadj : nullptr, st.f_muadj, st.f_sigadj, seed, 0, nullptr);
            if (rc) return rc;
            // the finish kernel leaves the value in the context scalar slot 0
            FADB_CUDA(cudaMemcpyAsync(st.mval, c.scalars + 0, sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
            if (guarded && !(flags & FADB_STRICT_SIGMA))
                return fadb_normal_lpdf_undo(st.f_shape, st.N, st.f_x, st.f_mu, st.f_sig, st.f_xadj, st.f_muadj, st.f_sigadj, seed, 0,
                                             st.d_red + 7);
            return FADB_OK;
        }
        if (st.fast == 4) {
            int rc = fadb_logpdf_async(st.f_op, st.f_shape, st.N, st.f_x, st.f_mu, st.f_sig, st.f_xadj, st.f_muadj,
                                       st.f_sigadj, seed, flags, st.d_red + 4);
            if (rc) return rc;
            FADB_CUDA(cudaMemcpyAsync(st.mval, st.d_red + 4, sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
            return FADB_OK;
        }
        if (st.fast == 3) {
            int rc = fadb_linreg_partial(st.f_rows, st.f_cols, st.f_X, st.f_x, st.f_w, st.f_part);
            if (rc) return rc;
            rc = fadb_linreg_finish(st.f_rows, st.f_cols, st.f_part, st.f_sig, st.f_wadj, st.f_sigadj, seed, 0, nullptr);
            if (rc) return rc;
            FADB_CUDA(cudaMemcpyAsync(st.mval, c.scalars + 2, sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
            return FADB_OK;
        }
        if (st.fast == 5) {
            int rc = fadb_linreg_partial(st.f_rows, st.f_cols, st.f_X, st.f_x, st.f_w, st.f_part);
            if (rc) return rc;
            return fadb_lsq_finish(st.f_cols, st.f_part, st.f_wadj, seed, 0, st.mval);
        }
        if (st.fast == 6) return run_glm(c, st, seed);
        return FADB_ERR_STATE;
    }

    // what = 1 feval, 2 beval, 3 autodiff
    int execute(Context& c, int what, double seed, const double* seed_vec)
    {
        const int last = (int)plan.size() - 1;
        if (what == 3 && !seed_vec && !blocks.empty() && g_additive_enabled.load()) {
            // additive root: every term is a root of its own with a seed known now (analyse_additive_root)
            Stage& rs = *plan[last];
            std::vector<double> seeds;
            if (root_term_seeds(rs, seed, seeds)) {
                for (const TermBlock& b : blocks) {
                    int rc = execute_range(c, 3, seeds[b.leaf], nullptr, b.first, b.last);
                    if (rc) return rc;
                }
                // the root itself: its value from the term values, and the adjoints of scalar leaves it reads directly
                return run_map(c, rs, M_FB, seed, nullptr, nullptr);
            }
        }
        return execute_range(c, what, seed, seed_vec, 0, last);
    }

    // the block plan[first..last] with plan[last] as its root
    int execute_range(Context& c, int what, double seed, const double* seed_vec, int first_stage, int last)
    {
        Stage& rs = *plan[last];
        // A log-pdf root swept with seed 0 leaves EVERYTHING below it untouched (normal.hpp:160, cauchy.hpp:151,
        // uniform.hpp / bernoulli.hpp alike): no root sweep and no producer sweeps either -- their boundary adjoints
        // would otherwise still hold the seeds of the previous call.
        if ((what & 2) && !seed_vec && seed == 0.0 && rs.type == 0 && rs.term >= T_NORMAL) {
            what &= 1;
            if (!what) return FADB_OK;
        }
        const bool fast = rs.fast != 0 && what == 3 && !seed_vec && (rs.fast != 6 || glm_usable(c, rs));
        const bool skip_producers = fast && (rs.fast == 3 || rs.fast == 5 || rs.fast == 6);   // linreg / least squares / one-pass map subsume their dot stage
        if ((what & 1) && !skip_producers) {
            for (int k = first_stage; k < last; ++k) {
                Stage& st = *plan[k];
                int rc = st.type != 0 ? run_dot(c, st, M_F, nullptr, 0) : run_map(c, st, M_F, 0, nullptr, nullptr);
                if (rc) return rc;
            }
        }
        if (fast) return run_fast_root(c, rs, seed);
        // root stage: fused when the seed suffices (no global quantity needed before the sweep)
        const bool fusable = rs.type != 0 || rs.term == T_NONE || rs.term == T_SUM ||
                             (rs.term == T_NORMAL && (!rs.sig_vec || (flags & FADB_TRUST_SIGMA))) ||
                             (rs.term > T_NORMAL && (!rs.needs_check || (flags & FADB_TRUST_SIGMA)));
        int rc = 0;
        if (what == 3 && fusable) {
            rc = rs.type != 0 ? run_dot(c, rs, M_FB, seed_vec, seed) : run_map(c, rs, M_FB, seed, seed_vec, nullptr);
        } else {
            if (what & 1) rc = rs.type != 0 ? run_dot(c, rs, M_F, nullptr, 0) : run_map(c, rs, M_F, 0, nullptr, nullptr);
            if (!rc && (what & 2)) {
                if (rs.type != 0) {
                    if (!seed_vec) {
                        int n = (int)rs.out_size;
                        fill_kernel<<<(n + 127) / 128, 128, 0, c.stream>>>(rs.madj, n, seed);
                        c.launches++;
                    }
                    rc = run_dot(c, rs, M_B, seed_vec, seed);
                } else rc = run_map(c, rs, M_B, seed, seed_vec, nullptr);
            }
        }
        if (rc) return rc;
        if (what & 2) {
            for (int k = last - 1; k >= first_stage; --k) {
                Stage& st = *plan[k];
                if (st.type != 0) rc = run_dot(c, st, M_B, nullptr, 0);
                else rc = run_map(c, st, M_B, 0, st.out_size > 1 ? st.madj : nullptr, st.out_size > 1 ? nullptr : st.madj);
                if (rc) return rc;
            }
        }
        return FADB_OK;
    }
};

// ======================================================================== C ABI
#define EXPR_REQUIRE(e, cond, msg)                          \
    do {                                                    \
        if (!(cond)) {                                      \
            set_error(std::string("fastad_b200: ") + msg);  \
            return FADB_ERR_ARG;                            \
        }                                                   \
    } while (0)

static double host_unary(int op, double x)
{   // eager folding of constant operands happens on the host, as in unary.hpp:182-184
    switch (op) {
    case FADB_NEG: return -x; case FADB_SIN: return std::sin(x); case FADB_COS: return std::cos(x);
    case FADB_TAN: return std::tan(x); case FADB_ASIN: return std::asin(x); case FADB_ACOS: return std::acos(x);
    case FADB_ATAN: return std::atan(x); case FADB_EXP: return std::exp(x); case FADB_LOG: return std::log(x);
    case FADB_SQRT: return std::sqrt(x); case FADB_ERF: return std::erf(x);
    case FADB_SIGMOID: return 1 / (1 + std::exp(-x)); case FADB_SINH: return std::sinh(x);
    case FADB_COSH: return std::cosh(x); case FADB_TANH: return std::tanh(x); case FADB_MOCK2X: return 2 * x;
    case FADB_SQUARE: return x * x;
    }
    return NAN;
}
static double host_binary(int op, double x, double y)
{
    switch (op) {
    case FADB_ADD: return x + y; case FADB_SUB: return x - y; case FADB_MUL: return x * y;
    case FADB_DIV: return x / y; case FADB_MOCKB: return x - 2 * y;
    case FADB_LT: return x < y; case FADB_LE: return x <= y; case FADB_GT: return x > y; case FADB_GE: return x >= y;
    case FADB_EQ: return x == y; case FADB_NE: return x != y; case FADB_AND: return x != 0 && y != 0;
    case FADB_OR: return x != 0 || y != 0;
    }
    return NAN;
}
static double host_pow(long n, double b)
{   // PowFunc, pow.hpp:20-56
    if (n < 0) return b == 0 ? INFINITY : host_pow(-n, 1. / b);
    if (n == 0) return 1.;
    if (n % 2 == 0) { double t = host_pow(n / 2, b); return t * t; }
    return b * host_pow(n - 1, b);
}
static bool valid_unary(int op) { return (op >= 0 && op < FADB_UNARY_COUNT) || op == FADB_MOCK2X; }
static bool valid_binary(int op) { return (op >= 0 && op < FADB_BINARY_COUNT) || op == FADB_MOCKB; }

extern "C" {

int fadb_expr_create(fadb_expr** out)
{
    EXPR_REQUIRE(out, out != nullptr, "fadb_expr_create: null out");
    *out = new fadb_expr();
    return FADB_OK;
}

void fadb_expr_destroy(fadb_expr* e)
{
    if (!e) return;
    if (!e->owned.empty()) {
        Context* c;
        if (current_context(&c) == FADB_OK) cudaStreamSynchronize(c->stream);
        for (void* p : e->owned) cudaFree(p);
    }
    delete e;
}

// Structural sharing of PURE nodes: the reference's expression objects hold their children by value, so a sub-expression
// used twice (eta = dot(X, w) + b in two terms of a likelihood) reaches the builder as two identical sub-trees.  An identical
// node -- same kind, same functor, same children -- is returned instead of a copy, as fadb_expr_var already does for views of the
// same storage: the planner then sees ONE dot node read twice and plans its stage once.  Nodes with an identity of their own
// (placeholder assignments, op=, glue) are never merged; elementwise nodes are re-lowered at every use anyway.
static int find_pure(fadb_expr* e, NodeKind kind, int fn, long pw, int c0, int c1)
{
    // index built lazily over the nodes added since the last lookup: O(log n) per node, whatever the graph size
    for (; e->pure_indexed < e->nodes.size(); ++e->pure_indexed) {
        const ENode& n = e->nodes[e->pure_indexed];
        if ((n.kind == N_UNARY || n.kind == N_BINARY || n.kind == N_POW || n.kind == N_DOT) && !n.ch.empty())
            e->pure_index.emplace(std::array<long, 5>{(long)n.kind, (long)n.fn, n.pw, (long)n.ch[0], n.ch.size() > 1 ? (long)n.ch[1] : -1L},
                                  (int)e->pure_indexed);
    }
    auto it = e->pure_index.find(std::array<long, 5>{(long)kind, (long)fn, pw, (long)c0, (long)c1});
    return it == e->pure_index.end() ? -1 : it->second;
}

int fadb_expr_var(fadb_expr* e, int shape, size_t rows, size_t cols, double* val, double* adj)
{
    EXPR_REQUIRE(e, e && shape >= 0 && shape <= 2, "fadb_expr_var: bad arguments");
    // VarView copies alias the same storage (var_view.hpp): one node per distinct view
    for (size_t k = 0; k < e->nodes.size(); ++k) {
        const ENode& n = e->nodes[k];
        if (n.kind == N_VAR && val && n.val == val && n.adj == adj && n.rows == rows && n.cols == cols) return (int)k;
    }
    ENode n; n.kind = N_VAR; n.shape = shape; n.rows = rows; n.cols = cols; n.val = val; n.adj = adj;
    return e->add(std::move(n));
}

int fadb_expr_const_scalar(fadb_expr* e, double c)
{
    EXPR_REQUIRE(e, e != nullptr, "fadb_expr_const_scalar: null expr");
    for (size_t k = 0; k < e->nodes.size(); ++k)
        if (e->nodes[k].kind == N_CONST_S && memcmp(&e->nodes[k].cval, &c, sizeof c) == 0) return (int)k;
    ENode n; n.kind = N_CONST_S; n.cval = c;
    return e->add(std::move(n));
}

int fadb_expr_const_view(fadb_expr* e, int shape, size_t rows, size_t cols, const double* val)
{
    EXPR_REQUIRE(e, e && shape >= 0 && shape <= 2, "fadb_expr_const_view: bad arguments");
    for (size_t k = 0; k < e->nodes.size(); ++k) {
        const ENode& m = e->nodes[k];
        if (m.kind == N_CONST_V && val && m.val == val && m.shape == shape && m.rows == rows && m.cols == cols) return (int)k;
    }
    ENode n; n.kind = N_CONST_V; n.shape = shape; n.rows = rows; n.cols = cols; n.val = val;
    return e->add(std::move(n));
}

int fadb_expr_rebind(fadb_expr* e, int leaf, double* val, double* adj)
{
    EXPR_REQUIRE(e, e && e->ok(leaf), "fadb_expr_rebind: bad leaf id");
    ENode& n = e->nodes[leaf];
    EXPR_REQUIRE(e, n.kind == N_VAR || n.kind == N_CONST_V, "fadb_expr_rebind: node is not a leaf");
    n.val = val; n.adj = adj;
    if (e->bound) {
        Context* c;
        int rc = current_context(&c);
        if (rc) return rc;
        return e->reupload_leaves(*c);
    }
    return FADB_OK;
}

int fadb_expr_unary(fadb_expr* e, int op, int child)
{
    EXPR_REQUIRE(e, e && e->ok(child) && valid_unary(op), "fadb_expr_unary: bad arguments");
    if (e->is_const_scalar(child)) return fadb_expr_const_scalar(e, host_unary(op, e->nodes[child].cval));
    if (int k = find_pure(e, N_UNARY, op, 0, child, -1); k >= 0) return k;
    const ENode& c = e->nodes[child];
    ENode n; n.kind = N_UNARY; n.fn = op; n.shape = c.shape; n.rows = c.rows; n.cols = c.cols; n.ch = {child};
    return e->add(std::move(n));
}

int fadb_expr_binary(fadb_expr* e, int op, int lhs, int rhs)
{
    EXPR_REQUIRE(e, e && e->ok(lhs) && e->ok(rhs) && valid_binary(op), "fadb_expr_binary: bad arguments");
    if (e->is_const_scalar(lhs) && e->is_const_scalar(rhs))      // binary.hpp:251-256
        return fadb_expr_const_scalar(e, host_binary(op, e->nodes[lhs].cval, e->nodes[rhs].cval));
    const ENode& l = e->nodes[lhs];
    const ENode& r = e->nodes[rhs];
    if (l.size() != 1 && r.size() != 1)
        EXPR_REQUIRE(e, l.rows == r.rows && l.cols == r.cols, "binary node: operand shapes differ (binary.hpp:86-90)");
    if (int k = find_pure(e, N_BINARY, op, 0, lhs, rhs); k >= 0) return k;
    const ENode& big = l.shape >= r.shape ? l : r;
    ENode n; n.kind = N_BINARY; n.fn = op; n.shape = big.shape;
    n.rows = std::max(l.rows, r.rows); n.cols = std::max(l.cols, r.cols); n.ch = {lhs, rhs};
    return e->add(std::move(n));
}

int fadb_expr_pow(fadb_expr* e, long pw, int child)
{
    EXPR_REQUIRE(e, e && e->ok(child) && pw > -32768 && pw < 32768, "fadb_expr_pow: exponent out of range");
    if (e->is_const_scalar(child)) return fadb_expr_const_scalar(e, host_pow(pw, e->nodes[child].cval));
    if (int k = find_pure(e, N_POW, 0, pw, child, -1); k >= 0) return k;
    const ENode& c = e->nodes[child];
    ENode n; n.kind = N_POW; n.pw = pw; n.shape = c.shape; n.rows = c.rows; n.cols = c.cols; n.ch = {child};
    return e->add(std::move(n));
}

static int make_reduce(fadb_expr* e, NodeKind k, int child)
{
    if (e->is_const_scalar(child)) return child;                 // sum.hpp:250-252
    ENode n; n.kind = k; n.ch = {child};
    return e->add(std::move(n));
}
int fadb_expr_sum(fadb_expr* e, int child)
{
    EXPR_REQUIRE(e, e && e->ok(child), "fadb_expr_sum: bad arguments");
    return make_reduce(e, N_SUM, child);
}
int fadb_expr_prod(fadb_expr* e, int child)
{
    EXPR_REQUIRE(e, e && e->ok(child), "fadb_expr_prod: bad arguments");
    return make_reduce(e, N_PROD, child);
}
static int make_iter(fadb_expr* e, NodeKind kind, int k, const int* ids)
{
    ENode n; n.kind = kind;
    for (int i = 0; i < k; ++i) {
        if (!e->ok(ids[i])) { set_error("fastad_b200: sum/prod over expressions: bad child id"); return FADB_ERR_ARG; }
        n.ch.push_back(ids[i]);
    }
    if (k > 0) { const ENode& c = e->nodes[ids[0]]; n.shape = c.shape; n.rows = c.rows; n.cols = c.cols; }
    return e->add(std::move(n));
}
int fadb_expr_sum_iter(fadb_expr* e, int k, const int* ids)
{
    EXPR_REQUIRE(e, e && k >= 0 && (k == 0 || ids), "fadb_expr_sum_iter: bad arguments");
    return make_iter(e, N_SUM_ITER, k, ids);
}
int fadb_expr_prod_iter(fadb_expr* e, int k, const int* ids)
{
    EXPR_REQUIRE(e, e && k >= 0 && (k == 0 || ids), "fadb_expr_prod_iter: bad arguments");
    return make_iter(e, N_PROD_ITER, k, ids);
}

int fadb_expr_dot(fadb_expr* e, int lhs, int rhs)
{
    EXPR_REQUIRE(e, e && e->ok(lhs) && e->ok(rhs), "fadb_expr_dot: bad arguments");
    const ENode& l = e->nodes[lhs];
    const ENode& r = e->nodes[rhs];
    EXPR_REQUIRE(e, l.cols == r.rows, "dot: lhs.cols() != rhs.rows() (dot.hpp:86)");
    if (int k = find_pure(e, N_DOT, 0, 0, lhs, rhs); k >= 0) return k;
    ENode n; n.kind = N_DOT; n.shape = r.shape == FADB_VEC ? FADB_VEC : FADB_MAT;
    n.rows = l.rows; n.cols = r.cols; n.ch = {lhs, rhs};
    return e->add(std::move(n));
}

int fadb_expr_normal(fadb_expr* e, int x, int mu, int sigma)
{
    EXPR_REQUIRE(e, e && e->ok(x) && e->ok(mu) && e->ok(sigma), "fadb_expr_normal: bad arguments");
    const size_t nx = e->nodes[x].size(), nm = e->nodes[mu].size(), ns = e->nodes[sigma].size();
    const bool dense = e->nodes[sigma].shape == FADB_MAT && e->nodes[sigma].rows == nx && e->nodes[sigma].cols == nx && nx > 1;
    EXPR_REQUIRE(e, (nm == 1 || nm == nx) && (ns == 1 || ns == nx || dense), "normal_adj_log_pdf: operand sizes differ");
    ENode n; n.kind = N_NORMAL; n.ch = {x, mu, sigma};
    return e->add(std::move(n));
}

int fadb_expr_eq(fadb_expr* e, int var, int child)
{
    EXPR_REQUIRE(e, e && e->ok(var) && e->ok(child) && e->nodes[var].kind == N_VAR, "fadb_expr_eq: target must be a Var");
    const ENode& w = e->nodes[var];
    EXPR_REQUIRE(e, w.size() == e->nodes[child].size(), "placeholder and expression shapes differ (eq.hpp:79-80)");
    ENode n; n.kind = N_EQ; n.shape = w.shape; n.rows = w.rows; n.cols = w.cols; n.ch = {var, child};
    return e->add(std::move(n));
}

int fadb_expr_glue(fadb_expr* e, int lhs, int rhs)
{
    EXPR_REQUIRE(e, e && e->ok(lhs) && e->ok(rhs), "fadb_expr_glue: bad arguments");
    const ENode& r = e->nodes[rhs];
    ENode n; n.kind = N_GLUE; n.shape = r.shape; n.rows = r.rows; n.cols = r.cols; n.ch = {lhs, rhs};
    return e->add(std::move(n));
}

int fadb_expr_logpdf(fadb_expr* e, int dist, int a, int b, int c)
{
    EXPR_REQUIRE(e, e && dist >= FADB_CAUCHY && dist <= FADB_BERNOULLI && e->ok(a) && e->ok(b) &&
                 (dist == FADB_BERNOULLI || e->ok(c)), "fadb_expr_logpdf: bad arguments");
    size_t N = std::max(e->nodes[a].size(), e->nodes[b].size());
    if (dist != FADB_BERNOULLI) N = std::max(N, e->nodes[c].size());
    for (int id : {a, b, dist == FADB_BERNOULLI ? a : c})
        EXPR_REQUIRE(e, e->nodes[id].size() == 1 || e->nodes[id].size() == N, "adjusted log-pdf: operand sizes differ");
    ENode n; n.kind = N_LOGPDF; n.fn = dist; n.ch = {a, b};
    if (dist != FADB_BERNOULLI) n.ch.push_back(c);
    return e->add(std::move(n));
}

int fadb_expr_det(fadb_expr* e, int policy, int take_log, int child)
{
    EXPR_REQUIRE(e, e && e->ok(child) && policy >= FADB_DET_FULLPIVLU && policy <= FADB_DET_LLT, "fadb_expr_det: bad arguments");
    const ENode& c = e->nodes[child];
    EXPR_REQUIRE(e, c.shape != FADB_SCL && c.rows == c.cols, "det: the operand must be a square matrix (det.hpp:33, 50)");
    ENode n; n.kind = N_DET; n.fn = policy; n.pw = take_log ? 1 : 0; n.ch = {child};
    return e->add(std::move(n));
}

int fadb_expr_wishart(fadb_expr* e, int x, int v, double df)
{
    EXPR_REQUIRE(e, e && e->ok(x) && e->ok(v), "fadb_expr_wishart: bad arguments");
    const ENode& a = e->nodes[x];
    const ENode& b = e->nodes[v];
    EXPR_REQUIRE(e, a.shape == FADB_MAT && b.shape == FADB_MAT && a.rows == a.cols && b.rows == b.cols && a.rows == b.rows,
                 "wishart_adj_log_pdf: X and V must be square matrices of one size (wishart.hpp:25-27)");
    ENode n; n.kind = N_WISHART; n.cval = df; n.ch = {x, v};
    return e->add(std::move(n));
}

int fadb_expr_norm(fadb_expr* e, int child)
{   // squaredNorm = sum of x*x with child seed 2*seed*x (norm.hpp:47-57)
    EXPR_REQUIRE(e, e && e->ok(child), "fadb_expr_norm: bad arguments");
    int sq = fadb_expr_unary(e, FADB_SQUARE, child);
    if (sq < 0) return sq;
    return fadb_expr_sum(e, sq);
}

int fadb_expr_transpose(fadb_expr* e, int child)
{
    EXPR_REQUIRE(e, e && e->ok(child), "fadb_expr_transpose: bad arguments");
    const ENode& c = e->nodes[child];
    EXPR_REQUIRE(e, c.shape != FADB_SCL, "transpose of a scalar (transpose.hpp:25)");
    ENode n; n.kind = N_TRANSPOSE; n.shape = FADB_MAT; n.rows = c.cols; n.cols = c.rows; n.ch = {child};
    return e->add(std::move(n));
}

int fadb_expr_if_else(fadb_expr* e, int cond, int if_node, int else_node)
{
    EXPR_REQUIRE(e, e && e->ok(cond) && e->ok(if_node) && e->ok(else_node), "fadb_expr_if_else: bad arguments");
    EXPR_REQUIRE(e, e->nodes[cond].size() == 1, "if_else: the condition must be a scalar (if_else.hpp:52-55)");
    const ENode& a = e->nodes[if_node];
    const ENode& b = e->nodes[else_node];
    EXPR_REQUIRE(e, a.rows == b.rows && a.cols == b.cols, "if_else: branch shapes differ (if_else.hpp:71-72)");
    if (e->is_const_scalar(cond)) return e->nodes[cond].cval != 0 ? if_node : else_node;
    ENode n; n.kind = N_IF_ELSE; n.shape = a.shape; n.rows = a.rows; n.cols = a.cols; n.ch = {cond, if_node, else_node};
    return e->add(std::move(n));
}

int fadb_expr_opeq(fadb_expr* e, int op, int var, int child)
{
    EXPR_REQUIRE(e, e && e->ok(var) && e->ok(child) && e->nodes[var].kind == N_VAR && op >= FADB_ADD && op <= FADB_DIV,
                 "fadb_expr_opeq: target must be a Var and op one of + - * /");
    const ENode& w = e->nodes[var];
    const ENode& c = e->nodes[child];
    EXPR_REQUIRE(e, c.size() == 1 || (w.rows == c.rows && w.cols == c.cols), "op=: shapes differ (eq.hpp:234-237)");
    ENode n; n.kind = N_OPEQ; n.fn = op; n.shape = w.shape; n.rows = w.rows; n.cols = w.cols; n.ch = {var, child};
    return e->add(std::move(n));
}

int fadb_expr_node_size(fadb_expr* e, int node, size_t* rows, size_t* cols)
{
    EXPR_REQUIRE(e, e && e->ok(node), "fadb_expr_node_size: bad node id");
    if (rows) *rows = e->nodes[node].rows;
    if (cols) *cols = e->nodes[node].cols;
    return FADB_OK;
}

int fadb_expr_is_const(fadb_expr* e, int node)
{
    if (!e || !e->ok(node)) return 0;
    NodeKind k = e->nodes[node].kind;
    return (k == N_CONST_S || k == N_CONST_V) ? 1 : 0;
}

/* host-only: cut the expression into stages (no CUDA call) */
int fadb_expr_plan(fadb_expr* e, int root, size_t out_cache_size[2])
{
    EXPR_REQUIRE(e, e && e->ok(root), "fadb_expr_plan: bad root id");
    int rc = e->do_plan(root);
    if (rc) { set_error("fastad_b200: cannot plan expression: " + e->err); return rc; }
    if (out_cache_size) { out_cache_size[0] = e->cache_vals; out_cache_size[1] = e->cache_adjs; }
    return FADB_OK;
}

// Text of the compile-time constants that specialise stage `stage` of the plan for `mode`
// (1 forward, 2 backward, 3 fused); host only.  FADB_ERR_UNSUPPORTED: the stage is not a
// specialisable elementwise program (control flow, non-finite immediates, a dot / dense stage).
int fadb_expr_jit_program(fadb_expr* e, int stage, int mode, char* buf, size_t cap)
{
    EXPR_REQUIRE(e, e && buf && cap > 0, "fadb_expr_jit_program: null argument");
    EXPR_REQUIRE(e, e->planned, "fadb_expr_jit_program: plan first");
    EXPR_REQUIRE(e, stage >= 0 && stage < (int)e->plan.size() && ((mode >= 1 && mode <= 3) || mode == 6 || (mode >= 9 && mode <= 11)),
                 "fadb_expr_jit_program: bad stage or mode");
    Stage& st = *e->plan[stage];
    std::string text;
    if (st.type != 0) { set_error("fastad_b200: not an elementwise stage"); return FADB_ERR_UNSUPPORTED; }
    if (st.enc.empty()) fadb_expr::encode_stage(st);
    if (mode == 6) {         // the one-pass dot -> map -> sum kernel of a fast == 6 root stage
        if (st.fast != 6 || !fadb_expr::glm_program_text(st, text)) { set_error("fastad_b200: not a one-pass (fast=6) stage"); return FADB_ERR_UNSUPPORTED; }
    } else if (!fadb_expr::jit_program_text(st, mode & 7, text, (mode & 8) != 0)) {      // mode | 8: the operand-ring form (JIT_RING)
        set_error("fastad_b200: stage is not specialisable"); return FADB_ERR_UNSUPPORTED;
    }
    if (text.size() + 1 > cap) { set_error("fastad_b200: buffer too small"); return FADB_ERR_ARG; }
    memcpy(buf, text.c_str(), text.size() + 1);
    return FADB_OK;
}

int fadb_glm_enable(int on) { return g_glm_enabled.exchange(on ? 1 : 0); }
int fadb_jit_ring_enable(int on) { return g_ring_enabled.exchange(on ? 1 : 0); }
int fadb_additive_root_enable(int on) { return g_additive_enabled.exchange(on ? 1 : 0); }

int fadb_expr_describe(fadb_expr* e, char* buf, size_t cap)
{
    EXPR_REQUIRE(e, e && e->planned && buf && cap > 0, "fadb_expr_describe: plan the expression first");
    std::string s = e->describe();
    size_t n = std::min(cap - 1, s.size());
    memcpy(buf, s.data(), n);
    buf[n] = 0;
    return (int)s.size();
}

int fadb_expr_bind(fadb_expr* e, int root, unsigned flags, size_t out_cache_size[2])
{
    int rc = fadb_expr_plan(e, root, out_cache_size);
    if (rc) return rc;
    e->flags = flags;
    Context* c;
    rc = current_context(&c);
    if (rc) return rc;
    return e->materialize(*c);
}

static int fetch_value(fadb_expr* e, Context& c, double* host_value)
{
    if (!host_value) return FADB_OK;
    Stage& rs = *e->plan.back();
    FADB_CUDA(cudaMemcpyAsync(host_value, rs.mval, rs.out_size * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    FADB_CUDA(cudaStreamSynchronize(c.stream));
    return FADB_OK;
}

static int run_expr(fadb_expr* e, int what, double seed, const double* seed_dev, double* host_value)
{
    EXPR_REQUIRE(e, e && e->bound, "expression is not bound (call fadb_expr_bind first)");
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    rc = e->execute(*c, what, seed, seed_dev);
    if (rc) return rc;
    return (what & 1) ? fetch_value(e, *c, host_value) : FADB_OK;
}

int fadb_expr_feval(fadb_expr* e, double* host_value) { return run_expr(e, 1, 0, nullptr, host_value); }
int fadb_expr_beval(fadb_expr* e, double seed, const double* seed_dev) { return run_expr(e, 2, seed, seed_dev, nullptr); }
int fadb_expr_autodiff(fadb_expr* e, double seed, const double* seed_dev, double* host_value)
{
    return run_expr(e, 3, seed, seed_dev, host_value);
}

int fadb_expr_value_ptr(fadb_expr* e, const double** dev_value, size_t* size)
{
    EXPR_REQUIRE(e, e && e->bound && dev_value, "fadb_expr_value_ptr: expression is not bound");
    *dev_value = e->plan.back()->mval;
    if (size) *size = e->plan.back()->out_size;
    return FADB_OK;
}

}  // extern "C"
