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
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <vector>

#include "common.cuh"
#include "functors.cuh"

extern "C" int fadb_sum_unary_async(int, size_t, const double*, double*, double, unsigned, double*);
extern "C" int fadb_sigma_check(size_t, const double*, double*);
extern "C" int fadb_normal_lpdf_partial(int, size_t, const double*, const double*, const double*, double*,
                                        double*, double*, double, unsigned, const double*, double*);
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

namespace fadb {

// ======================================================================== device program
enum : int { I_LEAF = 0, I_CONST, I_UNARY, I_BINARY, I_POW, I_EQ, I_GLUE, I_JZ, I_JMP, I_PHI, I_OPEQ };
enum : int { T_NONE = 0, T_SUM, T_PROD, T_NORMAL, T_CAUCHY, T_UNIFORM, T_BERNOULLI };
enum : int { M_F = 1, M_B = 2, M_FB = 3 };

struct Ins {
    int op, fn, a, b, leaf, pad;
    double imm;
};
// device encoding: 16 bytes (one LDS.128), a single flat opcode so that the interpreter has ONE
// dense switch (a jump table) per direction instead of nested compare chains
enum : short { F_LEAF_V = 0, F_LEAF_S, F_CONST, F_EQ, F_GLUE, F_POW, F_POW2, F_UNARY = 8 /* +fn, mock2x = +16 */,
               F_BINARY = 26 /* +fn, mockb = +12 */, F_JZ = 40, F_JMP = 41, F_PHI = 42, F_OPEQ = 44 /* +op */ };
struct DIns {
    short code, a, b, leaf;   // POW: exponent in `leaf`
    double imm;
};
struct LeafDesc {
    const double* val;
    double* adj;
    int scalar;      // 1: one value broadcast over the stage's elements
    int adj_mode;    // 0 none (constant), 1 accumulate (+=, var_view.hpp:93), 2 assign (stage boundary)
    int channel;     // reduction channel of a broadcast scalar's adjoint, or -1
    int assigned;    // 1: an I_EQ in this program owns the adjoint write-back
};

constexpr int MAXS = 64;    // SSA slots per stage program
constexpr int MAXL = 24;    // distinct leaves per stage
constexpr int MAXCH = 8;    // reduction channels per stage

struct StageArgs {
    const DIns* prog;
    const LeafDesc* leaves;
    int nins, nleaves;
    unsigned long long N;
    int mode, term, root;
    int nx, nm, ns;            // T_NORMAL operand slots
    int mu_vec, sig_vec;       // operand is per-element (1) or one broadcast scalar (0)
    const double* sig_ptr;     // scalar sigma value (device) when !sig_vec
    double* out_vec;           // per-element value store of the stage root (may be null)
    double* out_scalar;        // reduced value of the stage (T_SUM / T_PROD / T_NORMAL)
    const double* seed_vec;    // per-element seed (vector-valued stage), or null
    const double* seed_ptr;    // device scalar seed (inner stage), or null
    double seed;               // host scalar seed (root stage)
    double* red;               // [8] stage scratch: T_PROD {pnz, zeros, val}; T_NORMAL [1] = #invalid
    const double* gate;        // if non-null and *gate != 0: adjoint sweep is a no-op (normal.hpp:457)
    int nch, term_ch;
    double* partials;
    unsigned int* ticket;
    double* scratch;           // BIG programs (scalar stages only): v | g | lacc in global memory
};

__device__ __forceinline__ double block_prod_all(double p, double* sm)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    p = warp_prod(p);
    if (lane == 0) sm[warp] = p;
    __syncthreads();
    double r = 1.0;
    if (warp == 0) r = warp_prod(lane < nwarp ? sm[lane] : 1.0);
    __syncthreads();
    return r;   // valid in thread 0
}

// BIG = false: program and leaf table staged in shared memory, node values in thread-local
// storage.  BIG = true (scalar stages with more than MAXS nodes, e.g. ad::sum over a thousand
// scalar Vars, test/integration/ad_inttest.cpp:143-156): one thread, everything in global memory.
// STORE selects where a thread keeps its node values / adjoints / leaf accumulators:
//   0 = thread-local arrays (fallback), 1 = dynamic shared memory, laid out [slot][thread] so
//   that a warp access is conflict-free (default: local arrays of this size spill through L1 into
//   L2/DRAM -- ncu showed 3.6x the algorithmic DRAM writes), 2 = global scratch (BIG programs).
template <int STORE>
__global__ void __launch_bounds__(128)
stage_kernel(StageArgs a)
{
    constexpr bool BIG = STORE == 2;
    extern __shared__ double s_dyn[];
    __shared__ DIns s_prog_sm[BIG ? 1 : MAXS];
    __shared__ LeafDesc s_leaf_sm[BIG ? 1 : MAXL];
    __shared__ double s_red[32 * MAXCH];
    __shared__ bool s_last;
    const DIns* s_prog = BIG ? a.prog : s_prog_sm;
    const LeafDesc* s_leaf = BIG ? a.leaves : s_leaf_sm;
    if (!BIG) {
        for (int k = threadIdx.x; k < a.nins; k += blockDim.x) s_prog_sm[k] = a.prog[k];
        for (int k = threadIdx.x; k < a.nleaves; k += blockDim.x) s_leaf_sm[k] = a.leaves[k];
        __syncthreads();
    }

    const bool fwd_out = a.mode & M_F;     // forward results are stored in this launch
    bool back = a.mode & M_B;
    double seed_s = a.seed_ptr ? *a.seed_ptr : a.seed;
    bool gated = a.gate && *a.gate != 0.0;
    double sig_s = 1.0;
    if (a.term == T_NORMAL && !a.sig_vec) {
        sig_s = a.sig_ptr ? *a.sig_ptr : 1.0;
        if (!(sig_s > 0)) gated = true;                       // normal.hpp:147, 231, 344
    }
    if (a.term >= T_NORMAL && !a.seed_vec && seed_s == 0.0) back = false;   // normal.hpp:160, cauchy.hpp:151
    double pr_pnz = 1.0, pr_zeros = 0.0, pr_val = 0.0;
    if (a.term == T_PROD && back) { pr_pnz = a.red[0]; pr_zeros = a.red[1]; pr_val = a.red[2]; }

    double acc[MAXCH];
#pragma unroll
    for (int c = 0; c < MAXCH; ++c) acc[c] = 0.0;
    double p_nz = 1.0, p_zeros = 0.0;

    double v_loc[STORE == 0 ? MAXS : 1], g_loc[STORE == 0 ? MAXS : 1], l_loc[STORE == 0 ? MAXL : 1];
    const int vs = STORE == 1 ? (int)blockDim.x : 1;             // slot stride
    double* vb = STORE == 0 ? v_loc : STORE == 1 ? s_dyn + threadIdx.x : a.scratch;
    double* gb = STORE == 0 ? g_loc : vb + (size_t)a.nins * vs;
    double* lb = STORE == 0 ? l_loc : gb + (size_t)a.nins * vs;
#define V(k) vb[(k) * vs]
#define G(k) gb[(k) * vs]
#define LA(k) lb[(k) * vs]
    const unsigned long long tid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const unsigned long long nth = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = tid; i < a.N; i += nth) {
        // ------------------------------------------------ forward (feval)
        for (int k = 0; k < a.nins; ++k) {
            const DIns I = s_prog[k];
            double r;
            switch (I.code) {
            case F_LEAF_V: r = s_leaf[I.leaf].val[i]; break;              // var_view.hpp:69
            case F_LEAF_S: r = s_leaf[I.leaf].val[0]; break;
            case F_CONST: r = I.imm; break;
            case F_EQ: {                                                   // eq.hpp:89-92
                r = V(I.a);
                if (fwd_out) { const LeafDesc& L = s_leaf[I.leaf]; const_cast<double*>(L.val)[L.scalar ? 0 : i] = r; }
                break;
            }
            case F_GLUE: r = V(I.b); break;                               // glue.hpp:67-71
            case F_POW: r = pow_int((long)I.leaf, V(I.a)); break;
            case F_POW2: { const double x = V(I.a); r = x * x; break; }
#define U(FN) case F_UNARY + FN: r = Unary<FN>::f(V(I.a)); break;
            U(FADB_NEG) U(FADB_SIN) U(FADB_COS) U(FADB_TAN) U(FADB_ASIN) U(FADB_ACOS) U(FADB_ATAN) U(FADB_EXP)
            U(FADB_LOG) U(FADB_SQRT) U(FADB_ERF) U(FADB_SIGMOID) U(FADB_SINH) U(FADB_COSH) U(FADB_TANH)
            U(FADB_SQUARE)
#undef U
            case F_UNARY + 16: r = 2 * V(I.a); break;                     // MockUnary
            case F_BINARY + FADB_ADD: r = V(I.a) + V(I.b); break;         // binary.hpp:265-338
            case F_BINARY + FADB_SUB: r = V(I.a) - V(I.b); break;
            case F_BINARY + FADB_MUL: r = V(I.a) * V(I.b); break;
            case F_BINARY + FADB_DIV: r = div_guarded(V(I.a), V(I.b)); break;
            case F_BINARY + FADB_LT: r = V(I.a) < V(I.b) ? 1.0 : 0.0; break;     // binary.hpp:347-470
            case F_BINARY + FADB_LE: r = V(I.a) <= V(I.b) ? 1.0 : 0.0; break;
            case F_BINARY + FADB_GT: r = V(I.a) > V(I.b) ? 1.0 : 0.0; break;
            case F_BINARY + FADB_GE: r = V(I.a) >= V(I.b) ? 1.0 : 0.0; break;
            case F_BINARY + FADB_EQ: r = V(I.a) == V(I.b) ? 1.0 : 0.0; break;
            case F_BINARY + FADB_NE: r = V(I.a) != V(I.b) ? 1.0 : 0.0; break;
            case F_BINARY + FADB_AND: r = (V(I.a) != 0 && V(I.b) != 0) ? 1.0 : 0.0; break;
            case F_BINARY + FADB_OR: r = (V(I.a) != 0 || V(I.b) != 0) ? 1.0 : 0.0; break;
            case F_BINARY + 12: r = V(I.a) - 2 * V(I.b); break;           // MockBinary
            // IfElseNode (if_else.hpp:64-68): the condition is a scalar, so the jump is uniform
            case F_JZ: r = 0.0; if (V(I.a) == 0.0) k = I.b - 1; break;
            case F_JMP: r = 0.0; k = I.b - 1; break;
            case F_PHI: r = V(I.leaf) != 0.0 ? V(I.a) : V(I.b); break;
            default: {                                                     // OpEqNode::feval, eq.hpp:240-247
                const double old = V(I.b), e = V(I.a);
                const int op = I.code - F_OPEQ;
                r = op == FADB_ADD ? old + e : op == FADB_SUB ? old - e : op == FADB_MUL ? old * e : old / e;
                if (fwd_out) { const LeafDesc& L = s_leaf[I.leaf]; const_cast<double*>(L.val)[i] = r; }
                break;
            }
            }
            V(k) = r;
        }
        // ------------------------------------------------ terminal, forward part
        double nrm_d = 0, nrm_s = 1, nrm_z = 0;
        if (a.term == T_SUM) acc[0] += V(a.root);
        else if (a.term == T_PROD) { if (V(a.root) == 0) p_zeros += 1.0; else p_nz *= V(a.root); }
        else if (a.term == T_NORMAL) {
            nrm_d = V(a.nx) - V(a.nm);
            nrm_s = V(a.ns);
            if (a.sig_vec) {
                if (!(nrm_s > 0)) acc[2] += 1.0;
                nrm_z = nrm_d / nrm_s;
                acc[0] += nrm_z * nrm_z;                       // normal.hpp:449, 546
                acc[1] += log(nrm_s);
            } else if (a.N == 1) {
                nrm_z = nrm_d / nrm_s;                         // sss, normal.hpp:153-155
                acc[0] += nrm_z * nrm_z;
            } else {
                acc[0] += nrm_d * nrm_d;                       // squaredNorm, normal.hpp:244
            }
        } else if (a.term == T_CAUCHY) {
            // CauchyAdjLogPDFNode (stat/cauchy.hpp): -log(gamma + d^2/gamma) per element, gamma > 0
            const double d = V(a.nx) - V(a.nm), gm = V(a.ns);
            if (!(gm > 0)) { acc[2] += 1.0; if (!a.sig_vec) gated = true; }
            acc[0] += a.N == 1 ? log(gm + (d * d) / gm)                 // cauchy.hpp:144-146
                               : log(gm + (1. / gm) * (d * d));         // cauchy.hpp:216-217
        } else if (a.term == T_UNIFORM) {
            // UniformAdjLogPDFNode (stat/uniform.hpp): -log(max - min), min < x < max
            const double x = V(a.nx), lo = V(a.nm), hi = V(a.ns);
            if (!(lo < x && x < hi)) acc[2] += 1.0;
            acc[0] += log(hi - lo);
        } else if (a.term == T_BERNOULLI) {
            // BernoulliAdjLogPDFNode (stat/bernoulli.hpp:119-141): x in {0,1}, 0 < p < 1
            const double x = V(a.nx), p = V(a.nm);
            const bool in_range = 0 < p && p < 1, zero_one = x == 0 || x == 1;
            if (!in_range || !zero_one) acc[2] += 1.0;
            if (!in_range) acc[0] += (p <= 0) ? (x == 0 ? 0.0 : -INFINITY) : (x == 1 ? 0.0 : -INFINITY);
            else acc[0] += x == 0 ? log(1 - p) : x == 1 ? log(p) : -INFINITY;
        } else if (fwd_out && a.out_vec) a.out_vec[i] = V(a.root);
        if (!back) continue;

        // ------------------------------------------------ adjoint sweep (beval)
        for (int k = 0; k < a.nins; ++k) G(k) = 0.0;
        for (int k = 0; k < a.nleaves; ++k) LA(k) = 0.0;
        const double sd = a.seed_vec ? a.seed_vec[i] : seed_s;
        if (!gated) {
            if (a.term == T_NORMAL) {
                const double x = V(a.nx), m = V(a.nm), s = nrm_s, d = nrm_d;
                if (a.sig_vec) {
                    const double s2 = s * s;
                    if (a.mu_vec) { G(a.nx) += sd * (m - x) / s2; G(a.nm) += sd * d / s2; }      // :563-564
                    else { G(a.nx) += (sd / s2) * (m - x); G(a.nm) += sd * (d / s2); }           // :471-474
                    G(a.ns) += (sd / s) * (nrm_z * nrm_z - 1.);                                  // :468, 560
                } else if (a.N == 1) {
                    const double inv_s = 1. / s, z = d * inv_s;                                  // :158-171
                    G(a.ns) += sd * (z * z - 1) * inv_s;
                    const double adj = sd * z * inv_s;
                    G(a.nm) += adj;
                    G(a.nx) += -adj;
                } else {
                    const double inv_s = 1. / s, inv_s_sq = inv_s * inv_s, c = sd * inv_s_sq;
                    G(a.nx) += c * (m - x);                                                      // :282, 371
                    G(a.nm) += a.mu_vec ? c * d : sd * (d * inv_s_sq);                           // :370, 277
                    G(a.ns) += sd * ((d * d) / (s * s) - 1.) * inv_s;                            // :273 per element
                }
            } else if (a.term == T_CAUCHY) {
                const double d = V(a.nx) - V(a.nm), gm = V(a.ns);
                if (a.N == 1) {                                                                   // cauchy.hpp:149-165
                    const double inner = gm + (d * d) / gm;
                    const double x0_adj = 2. * d / (gm * inner);
                    G(a.ns) += sd * (1. / gm * (x0_adj * d - 1));
                    G(a.nm) += sd * x0_adj;
                    G(a.nx) += sd * -x0_adj;
                } else {   // vector cases exactly as written in cauchy.hpp:220-237 (seed enters dx0, dgamma twice)
                    const double dx = (-2. * sd) * d / (gm * gm + d * d);
                    G(a.ns) += (-sd / gm) * (dx * d + 1);
                    G(a.nm) += (-sd) * dx;
                    G(a.nx) += dx;
                }
            } else if (a.term == T_UNIFORM) {                                                     // uniform.hpp:157-163
                const double adj = sd / (V(a.ns) - V(a.nm));
                G(a.ns) += -adj;
                G(a.nm) += adj;
            } else if (a.term == T_BERNOULLI) {                                                   // bernoulli.hpp:143-149
                const double x = V(a.nx), p = V(a.nm);
                G(a.nm) += x == 0 ? -sd / (1 - p) : sd / p;
            } else if (a.term == T_PROD) {
                const double e = V(a.root);                                                      // prod.hpp:194-207
                G(a.root) = (e == 0) ? sd * ((pr_zeros > 1) ? 0.0 : pr_pnz) : sd * (pr_val / e);
            } else {
                G(a.root) = sd;
            }
            for (int k = a.nins - 1; k >= 0; --k) {
                const DIns I = s_prog[k];
                const double s = G(k);
                switch (I.code) {
                case F_LEAF_V: case F_LEAF_S: LA(I.leaf) += s; break;
                case F_CONST: break;
                case F_EQ: {                                                                      // eq.hpp:101-107
                    const LeafDesc& L = s_leaf[I.leaf];
                    if (L.scalar && a.N > 1) { G(I.a) += s + LA(I.leaf); break; }                // not produced by the planner
                    const double tot = (L.adj[i] + LA(I.leaf)) + s;
                    L.adj[i] = tot;
                    LA(I.leaf) = 0.0;
                    G(I.a) += tot;
                    break;
                }
                case F_GLUE: G(I.b) += s; break;                                                  // lhs gets 0, glue.hpp:81-86
                case F_POW: G(I.a) += pow_bmap((long)I.leaf, s, V(I.a), V(k)); break;            // pow.hpp:101-162
                case F_POW2: { const double x = V(I.a); G(I.a) += 2. * s * (x == 0 ? 0 : V(k) / x); break; }
#define U(FN) case F_UNARY + FN: G(I.a) += Unary<FN>::b(s, V(I.a), V(k)); break;                 /* unary.hpp:81-89 */
                U(FADB_NEG) U(FADB_SIN) U(FADB_COS) U(FADB_TAN) U(FADB_ASIN) U(FADB_ACOS) U(FADB_ATAN) U(FADB_EXP)
                U(FADB_LOG) U(FADB_SQRT) U(FADB_ERF) U(FADB_SIGMOID) U(FADB_SINH) U(FADB_COSH) U(FADB_TANH)
                U(FADB_SQUARE)
#undef U
                case F_UNARY + 16: G(I.a) += s * 2; break;
                // binary.hpp:120-136: rhs seed first, then lhs
                case F_BINARY + FADB_ADD: G(I.b) += s; G(I.a) += s; break;
                case F_BINARY + FADB_SUB: G(I.b) += -s; G(I.a) += s; break;
                case F_BINARY + FADB_MUL: { const double x = V(I.a), y = V(I.b); G(I.b) += s * x; G(I.a) += s * y; break; }
                case F_BINARY + FADB_DIV: { const double y = V(I.b); G(I.b) += div_guarded(-s * V(k), y); G(I.a) += div_guarded(s, y); break; }
                case F_BINARY + 12: G(I.b) += -2. * s; G(I.a) += s; break;                        // MockBinary
                // comparisons: no adjoint (binary.hpp:124)
                case F_BINARY + FADB_LT: case F_BINARY + FADB_LE: case F_BINARY + FADB_GT: case F_BINARY + FADB_GE:
                case F_BINARY + FADB_EQ: case F_BINARY + FADB_NE: case F_BINARY + FADB_AND: case F_BINARY + FADB_OR: break;
                // IfElseNode::beval (if_else.hpp:70-78): only the taken branch is swept
                case F_PHI:
                    if (V(I.leaf) != 0.0) { G(I.a) += s; k = (int)I.imm + 1; }    // skip the else branch
                    else G(I.b) += s;
                    break;
                case F_JMP: if (V(I.a) == 0.0) k = I.leaf + 1; break;             // else was taken: skip the if branch
                case F_JZ: break;
                default: {                                                         // OpEqNode::beval, eq.hpp:249-271
                    const LeafDesc& L = s_leaf[I.leaf];
                    const double tot = (L.adj[i] + LA(I.leaf)) + s;
                    LA(I.leaf) = 0.0;
                    const double old = V(I.b), e = V(I.a);
                    const int op = I.code - F_OPEQ;
                    double ls, rs;
                    if (op == FADB_ADD) { ls = tot; rs = tot; }
                    else if (op == FADB_SUB) { ls = tot; rs = -tot; }
                    else if (op == FADB_MUL) { ls = tot * e; rs = tot * old; }
                    else { ls = tot / e; rs = -tot * old / (e * e); }
                    G(I.a) += rs;
                    L.adj[i] = ls;                                                 // reset_adj(); beval(lseed)
                    const_cast<double*>(L.val)[i] = old;                           // the previous value is restored
                    break;
                }
                }
            }
        }
        // ------------------------------------------------ leaf write-back: adj += seed (var_view.hpp:91-94)
        for (int k = 0; k < a.nleaves; ++k) {
            const LeafDesc& L = s_leaf[k];
            if (L.adj_mode == 0) continue;
            if (L.assigned) {          // reads that precede the assignment in program order
                if (LA(k) != 0.0) L.adj[i] += LA(k);
                continue;
            }
            if (L.scalar && a.N > 1) { acc[L.channel] += LA(k); continue; }
            if (L.adj_mode == 2) L.adj[i] = LA(k);
            else if (!gated) L.adj[i] += LA(k);
        }
    }

    // ---------------------------------------------------- grid reduction + scalar epilogue
    if (a.nch == 0 && a.term != T_PROD) return;
    if (a.term == T_PROD && (a.mode & M_F)) {
        const double bp = block_prod_all(p_nz, s_red);
        acc[0] = p_zeros;
        block_sum<MAXCH>(acc, s_red);
        if (threadIdx.x == 0) {
            a.partials[(size_t)blockIdx.x * (MAXCH + 1) + MAXCH] = bp;
            for (int c = 0; c < MAXCH; ++c) a.partials[(size_t)blockIdx.x * (MAXCH + 1) + c] = acc[c];
            __threadfence();
            s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        double p = 1.0, z = 0.0;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
            p *= __ldcg(&a.partials[(size_t)b * (MAXCH + 1) + MAXCH]);
            z += __ldcg(&a.partials[(size_t)b * (MAXCH + 1)]);
        }
        p = block_prod_all(p, s_red);
        double zz[1] = {z};
        block_sum<1>(zz, s_red);
        if (threadIdx.x == 0) {
            a.red[0] = p; a.red[1] = zz[0]; a.red[2] = (zz[0] > 0) ? 0.0 : p;                    // prod.hpp:178
            a.out_scalar[0] = a.red[2];
            *a.ticket = 0u;
        }
        return;
    }
    block_sum<MAXCH>(acc, s_red);
    if (threadIdx.x == 0) {
        for (int c = 0; c < MAXCH; ++c) a.partials[(size_t)blockIdx.x * (MAXCH + 1) + c] = acc[c];
        __threadfence();
        s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double tot[MAXCH];
#pragma unroll
    for (int c = 0; c < MAXCH; ++c) tot[c] = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x)
#pragma unroll
        for (int c = 0; c < MAXCH; ++c) tot[c] += __ldcg(&a.partials[(size_t)b * (MAXCH + 1) + c]);
    block_sum<MAXCH>(tot, s_red);
    if (threadIdx.x != 0) return;
    *a.ticket = 0u;
    if (a.mode & M_F) {
        if (a.term == T_SUM) a.out_scalar[0] = tot[0];
        else if (a.term == T_NORMAL) {
            double val;
            if (a.sig_vec) {
                a.red[1] = tot[2];
                val = (tot[2] != 0.0) ? -INFINITY : -0.5 * tot[0] - tot[1];                      // normal.hpp:451, 548
            } else {
                a.red[1] = (sig_s > 0) ? 0.0 : 1.0;
                if (!(sig_s > 0)) val = -INFINITY;
                else if (a.N == 1) val = -0.5 * tot[0] - log(sig_s);
                else val = -0.5 * (tot[0] / (sig_s * sig_s)) - (double)a.N * log(sig_s);         // normal.hpp:245-246
            }
            a.out_scalar[0] = val;
        } else if (a.term == T_CAUCHY || a.term == T_UNIFORM) {
            a.red[1] = tot[2];
            a.out_scalar[0] = tot[2] != 0.0 ? -INFINITY : -tot[0];
        } else if (a.term == T_BERNOULLI) {
            a.red[1] = tot[2];
            a.out_scalar[0] = tot[0];
        }
    }
    if (a.mode & M_B) {
        for (int k = 0; k < a.nleaves; ++k) {
            const LeafDesc& L = s_leaf[k];
            if (L.channel < 0 || L.adj_mode == 0) continue;
            const bool skip = gated || !back || (a.term > T_NORMAL && (a.mode & M_F) && tot[2] != 0.0);
            if (L.adj_mode == 2) L.adj[0] = skip ? 0.0 : tot[L.channel];
            else if (!skip) L.adj[0] += tot[L.channel];
        }
    }
}

#undef V
#undef G
#undef LA

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

struct Stage {
    int type = 0;                  // 0 = MAP (interpreter), 1 = DOT, 2 = TRANSPOSE, 3 = dense-covariance normal,
                                   // 4 = det / log_det (k = take_log, p = policy), 5 = wishart (sig_const = df)
    size_t N = 1;
    std::vector<Ins> prog;
    std::vector<LeafDesc> leaves;
    std::vector<int> leaf_node;    // node id of each leaf (-1: stage boundary)
    std::vector<int> leaf_stage;   // producing stage of a boundary leaf (-1 otherwise)
    std::map<int, int> eq_slot;    // placeholder node id -> slot holding its current value
    std::set<int> assigned;        // placeholder node ids assigned in this stage
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
    LeafDesc* d_leaves = nullptr;
    double* d_red = nullptr;       // [8]
    double* d_sig = nullptr;       // device copy of a constant scalar sigma
    bool big = false;              // scalar stage too large for the on-chip interpreter
    double* d_scratch = nullptr;
    int fast = 0;                  // 0 interpreter, 1 sum_unary, 2 normal leaves, 3 linreg, 4 cauchy/uniform/bernoulli leaves
    // fast-path operands
    int f_op = 0; const double* f_x = nullptr; double* f_xadj = nullptr;
    const double *f_mu = nullptr, *f_sig = nullptr; double *f_muadj = nullptr, *f_sigadj = nullptr;
    int f_shape = 0; const double* f_X = nullptr; const double* f_w = nullptr; double* f_wadj = nullptr;
    size_t f_rows = 0, f_cols = 0; double* f_part = nullptr;
};

}  // namespace fadb

using namespace fadb;

struct fadb_expr {
    std::vector<ENode> nodes;
    std::vector<std::unique_ptr<Stage>> plan;
    std::vector<void*> owned;        // device allocations owned by the expression
    int root = -1;
    bool planned = false, bound = false;
    unsigned flags = 0;
    std::string err;
    size_t cache_vals = 0, cache_adjs = 0;

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
            int sidx = plan_stage(nid);
            if (sidx < 0) return -1;
            Stage& ps = *plan[sidx];
            const bool sc = ps.out_size == 1 && st.N > 1;
            if (ps.out_size != st.N && !sc) { err = "shape mismatch at a stage boundary"; return -1; }
            int lf = leaf_index(st, -1, nullptr, nullptr, sc, 2, sidx);
            st.leaves[lf].adj_mode = 2;       // pointers are patched at bind time
            return emit(st, Ins{I_LEAF, 0, 0, 0, lf, 0, 0.0});
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
            int a = lower(n.ch[1], st); if (a < 0) return -1;
            int jmp = emit(st, Ins{I_JMP, 0, c, 0, jz, 0, 0.0});
            st.prog[jz].b = (int)st.prog.size();
            int b = lower(n.ch[2], st); if (b < 0) return -1;
            st.prog[jmp].b = (int)st.prog.size();
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
            err = "a fused vector stage is limited to " + std::to_string(MAXS) + " nodes and " +
                  std::to_string(MAXL) + " distinct leaves";
            return false;
        }
        st.term_ch = st.term == T_SUM ? 1 : st.term >= T_NORMAL ? 3 : st.term == T_PROD ? 1 : 0;
        int ch = st.term_ch;
        for (auto& L : st.leaves)
            if (L.scalar && L.adj_mode != 0 && st.N > 1) L.channel = ch++;
        if (ch > MAXCH) { err = "expression stage needs more than " + std::to_string(MAXCH) + " reduction channels"; return false; }
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
            st->m = (int)l.rows; st->k = (int)l.cols; st->p = (int)r.cols;
            st->out_size = n.size();
            st->L = operand(n.ch[0]); if (!err.empty()) return -1;
            st->R = operand(n.ch[1]); if (!err.empty()) return -1;
        } else if (n.kind == N_TRANSPOSE) {
            ENode& x = nodes[n.ch[0]];
            st->type = 2;
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

    void detect_fast_paths()
    {
        if (plan.empty()) return;
        Stage& st = *plan.back();
        st.fast = 0;
        if (st.type != 0 || plan.size() > 2) return;
        int lf;
        if (plan.size() == 2 && st.term != T_NORMAL) return;
        if (plan.size() == 1 && st.term == T_SUM && st.prog.size() == 2 && st.root == 1 && leaf_only(st, 0, &lf) &&
            !st.leaves[lf].scalar && (st.prog[1].op == I_UNARY || st.prog[1].op == I_POW) && st.prog[1].a == 0) {
            st.fast = 1;
            st.f_op = st.prog[1].op == I_UNARY ? st.prog[1].fn : 1000 + st.prog[1].fn;
            st.f_x = st.leaves[lf].val; st.f_xadj = st.leaves[lf].adj_mode ? st.leaves[lf].adj : nullptr;
            return;
        }
        if (st.term > T_NORMAL && plan.size() == 1) {
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
            if (plan.size() == 1 && x_leaf && m_leaf && s_leaf && st.prog.size() == 3) {
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
                    ds.k <= 4096 && plan.size() == 2) {
                    st.fast = 3;
                    st.f_X = ds.L.val; st.f_w = ds.R.val; st.f_wadj = ds.R.adj_mode ? ds.R.adj : nullptr;
                    st.f_rows = ds.m; st.f_cols = ds.k;
                    st.f_x = st.leaves[lx].val;
                    st.f_sig = st.leaves[ls].val; st.f_sigadj = st.leaves[ls].adj_mode ? st.leaves[ls].adj : nullptr;
                }
            }
        }
    }

    std::string describe() const
    {
        static const char* opn[] = {"leaf", "const", "unary", "binary", "pow", "eq", "glue", "jz", "jmp", "phi", "opeq"};
        static const char* tn[] = {"store", "sum", "prod", "normal", "cauchy", "uniform", "bernoulli"};
        std::ostringstream o;
        o << "stages=" << plan.size() << " cache={" << cache_vals << "," << cache_adjs << "}\n";
        for (size_t k = 0; k < plan.size(); ++k) {
            const Stage& st = *plan[k];
            if (st.type == 5) { o << "stage " << k << ": wishart p=" << st.m << " df=" << st.sig_const << "\n"; continue; }
            if (st.type == 4) { o << "stage " << k << ": " << (st.k ? "log_det" : "det") << " n=" << st.m << " policy=" << st.p << "\n"; continue; }
            if (st.type == 3) { o << "stage " << k << ": dense normal n=" << st.m << (st.k ? " (vvm)" : " (vsm)") << "\n"; continue; }
            if (st.type == 2) { o << "stage " << k << ": transpose " << st.m << "x" << st.k << "\n"; continue; }
            if (st.type == 1) { o << "stage " << k << ": dot " << st.m << "x" << st.k << " * " << st.k << "x" << st.p << "\n"; continue; }
            o << "stage " << k << ": map N=" << st.N << " term=" << tn[st.term] << " ins=" << st.prog.size()
              << " leaves=" << st.leaves.size() << " channels=" << st.nch << " fast=" << st.fast
              << (st.big ? " big" : "") << "\n";
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
            std::vector<DIns> enc(st.prog.size());
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
            FADB_CUDA(cudaMemcpyAsync(st.d_prog, enc.data(), enc.size() * sizeof(DIns), cudaMemcpyHostToDevice, c.stream));
            FADB_CUDA(cudaStreamSynchronize(c.stream));     // `enc` is a temporary
            FADB_CUDA(cudaMemcpyAsync(st.d_leaves, st.leaves.data(), st.leaves.size() * sizeof(LeafDesc), cudaMemcpyHostToDevice, c.stream));
            if (st.term == T_NORMAL && st.sig_is_const) {
                if ((rc = dev_alloc(sizeof(double), (void**)&st.d_sig))) return rc;
                FADB_CUDA(cudaMemcpyAsync(st.d_sig, &st.sig_const, sizeof(double), cudaMemcpyHostToDevice, c.stream));
            }
            if (st.fast == 3) {
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
        }
        FADB_CUDA(cudaStreamSynchronize(c.stream));
        detect_fast_paths();
        return FADB_OK;
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
        if (st.big) {
            stage_kernel<2><<<1, 32, 0, c.stream>>>(a);
            c.launches++;
            FADB_CUDA(cudaGetLastError());
            return FADB_OK;
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
            static const Stage* owner[kMaxDevices] = {};
            double sd = (mode & M_B) ? seed : 0.0;
            if (mode == M_B && !seed_vec) {
                FADB_CUDA(cudaMemcpyAsync(&sd, st.madj, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
                FADB_CUDA(cudaStreamSynchronize(c.stream));
            }
            const bool mine = owner[c.device] == &st;
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
            owner[c.device] = &st;
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
            int n = std::max(st.m * st.k, st.k * st.p);
            dot_bwd_kernel<<<(n + threads - 1) / threads, threads, 0, c.stream>>>(
                st.m, st.k, st.p, st.L.val, st.R.val, A, st.L.adj_mode ? st.L.adj : nullptr, st.L.adj_mode,
                st.R.adj_mode ? st.R.adj : nullptr, st.R.adj_mode);
            c.launches++;
        }
        FADB_CUDA(cudaGetLastError());
        return FADB_OK;
    }

    int run_fast_root(Context& c, Stage& st, double seed)
    {
        if (st.fast == 1)
            return fadb_sum_unary_async(st.f_op, st.N, st.f_x, st.f_xadj, seed, 0, st.mval);
        if (st.fast == 2) {
            const double* inv = nullptr;
            if (st.sig_vec && seed != 0.0 && (st.f_xadj || st.f_muadj || st.f_sigadj) && !(flags & FADB_TRUST_SIGMA)) {
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
        return FADB_ERR_STATE;
    }

    // what = 1 feval, 2 beval, 3 autodiff
    int execute(Context& c, int what, double seed, const double* seed_vec)
    {
        const int last = (int)plan.size() - 1;
        Stage& rs = *plan[last];
        const bool fast = rs.fast != 0 && what == 3 && !seed_vec;
        const size_t first = (fast && rs.fast == 3) ? plan.size() : 0;   // linreg subsumes its dot stage
        if (what & 1) {
            for (size_t k = first; k + 1 < plan.size(); ++k) {
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
            for (int k = last - 1; k >= 0; --k) {
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
    ENode n; n.kind = N_CONST_S; n.cval = c;
    return e->add(std::move(n));
}

int fadb_expr_const_view(fadb_expr* e, int shape, size_t rows, size_t cols, const double* val)
{
    EXPR_REQUIRE(e, e && shape >= 0 && shape <= 2, "fadb_expr_const_view: bad arguments");
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
    const ENode& big = l.shape >= r.shape ? l : r;
    ENode n; n.kind = N_BINARY; n.fn = op; n.shape = big.shape;
    n.rows = std::max(l.rows, r.rows); n.cols = std::max(l.cols, r.cols); n.ch = {lhs, rhs};
    return e->add(std::move(n));
}

int fadb_expr_pow(fadb_expr* e, long pw, int child)
{
    EXPR_REQUIRE(e, e && e->ok(child) && pw > -32768 && pw < 32768, "fadb_expr_pow: exponent out of range");
    if (e->is_const_scalar(child)) return fadb_expr_const_scalar(e, host_pow(pw, e->nodes[child].cval));
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
