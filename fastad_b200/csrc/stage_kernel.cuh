// stage_kernel.cuh — the stage kernel of the generic expression path (see expr.cu for the planner).
//
// ONE source, two builds:
//  * nvcc, inside libfastad_b200.so: the tape INTERPRETER.  The stage program (16-byte instructions)
//    is data; each thread walks it with one dense switch per direction, node values in local arrays.
//  * NVRTC at bind time (jit.cu), FADB_JIT defined: the SPECIALISED kernel of one stage program.  The
//    program, the leaf attributes and the terminal configuration are compile-time constants
//    (fadb_jit_program, generated from the planner's output), both instruction loops are fully
//    unrolled, every switch folds to its one case and node values live in registers: the same code
//    the type-fused path (include/fastad_b200/fused.cuh) gets from the expression TYPE under nvcc,
//    here for host code built with plain g++.  Leaf POINTERS, N and the seed stay kernel arguments,
//    so rebinding or resizing does not recompile.
// Everything in this file must stay NVRTC-clean: no host or standard-library includes.
#pragma once
#ifndef __CUDACC_RTC__
#include "../../include/fastad_b200/device_math.cuh"
#include "device_reduce.cuh"
#endif

namespace fadb {

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
// The three limits above are the interpreter's (fixed-size local and shared arrays).  A specialised
// kernel sizes everything from its own program, so stages up to the limits below are accepted and
// simply have no interpreter form (they need libnvrtc at run time):
constexpr int MAXS_JIT = 512, MAXL_JIT = 64, MAXCH_JIT = 32;

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
    int vec_ok;                // specialised vector path: every per-element leaf pointer is 32-byte aligned
    // leaf pointers again, in PARAMETER space: the specialised kernel indexes them with compile-time
    // constants, so they are constant-bank operands (no registers, no loads); the interpreter uses `leaves`
    const double* lp_val[MAXL_JIT];
    double* lp_adj[MAXL_JIT];
};

#ifdef FADB_JIT
#ifndef FADB_JIT_PF_DIST
#define FADB_JIT_PF_DIST 1   // trips ahead the operand prefetch runs
#endif
struct JitLeaf { int scalar, adj_mode, channel, assigned; };   // the compile-time part of a LeafDesc
#endif
//@@FADB_JIT_PROGRAM@@  (jit.cu splices the generated program constants in here)

#ifdef FADB_JIT
#define K_NINS JIT_NINS
#define K_NLEAVES JIT_NLEAVES
#define K_INS(k) JIT_PROG[k]
#ifndef JIT_VEC
#define JIT_VEC 1
#endif
#ifndef JIT_HAS_CF
#define JIT_HAS_CF 0
#endif
#ifndef JIT_RING
#define JIT_RING 0
#endif
#define K_LVAL(k) a.lp_val[k]
#define K_LADJ(k) a.lp_adj[k]
#define K_LEAF_SCALAR(k) JIT_LEAF[k].scalar
#define K_LEAF_ADJ_MODE(k) JIT_LEAF[k].adj_mode
#define K_LEAF_CHANNEL(k) JIT_LEAF[k].channel
#define K_LEAF_ASSIGNED(k) JIT_LEAF[k].assigned
#define K_MODE JIT_MODE
#define K_TERM JIT_TERM
#define K_ROOT JIT_ROOT
#define K_NX JIT_NX
#define K_NM JIT_NM
#define K_NS JIT_NS
#define K_MU_VEC JIT_MU_VEC
#define K_SIG_VEC JIT_SIG_VEC
#define K_NCH JIT_NCH
#define K_SLOTS JIT_NINS
#define K_ACC (JIT_NCH > 3 ? JIT_NCH : 3)      // accumulators: the terminal's own channels are 0..2
#define K_LSLOTS (JIT_NLEAVES > 0 ? JIT_NLEAVES : 1)
#define K_UNROLL _Pragma("unroll")
#if JIT_VEC > 1
#define K_LOADV(k) lv[k][e]
#else
#define K_LOADV(k) K_LVAL(k)[i]
#endif
#else
#define K_NINS a.nins
#define K_NLEAVES a.nleaves
#define K_INS(k) s_prog[k]
#define K_LVAL(k) s_leaf[k].val
#define K_LADJ(k) s_leaf[k].adj
#define K_LEAF_SCALAR(k) s_leaf[k].scalar
#define K_LEAF_ADJ_MODE(k) s_leaf[k].adj_mode
#define K_LEAF_CHANNEL(k) s_leaf[k].channel
#define K_LEAF_ASSIGNED(k) s_leaf[k].assigned
#define K_MODE a.mode
#define K_TERM a.term
#define K_ROOT a.root
#define K_NX a.nx
#define K_NM a.nm
#define K_NS a.ns
#define K_MU_VEC a.mu_vec
#define K_SIG_VEC a.sig_vec
#define K_NCH a.nch
#define K_SLOTS MAXS
#define K_ACC MAXCH
#define K_LSLOTS MAXL
#define K_UNROLL
#define K_LOADV(k) K_LVAL(k)[i]
#endif

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
#if defined(FADB_JIT) && JIT_RING
// Operand ring of the one-element-per-thread form (stages with placeholder stores).  ncu on the Black-Scholes call
// expression: half of all stall samples sat on TEN serialised global loads per element -- the operands at their first
// use, every re-read of a placeholder behind its store (the compiler must assume the store aliases the other leaves), the
// old adjoints of the three placeholders and of the three variables -- and `prefetch.global.L1` does not shorten them (L1 hit
// rate 27 %).  Loading everything up front costs registers the 64-register budget does not have.  So the loads become
// cp.async copies (global -> shared, no registers) issued ONE TRIP AHEAD into a per-thread slot column: values are copied
// after the forward sweep of the previous trip has consumed its own, old adjoints (and the per-element seed) after its
// write-back; a trip then waits on data that has been in flight for half a trip or more.  A placeholder read behind its
// assignment takes the register of the assigning instruction.  Each thread reads only what it copied itself: no barriers.
// The planner only emits JIT_RING for programs without op= / if_else, and the launch falls back to the plain kernel when
// any written operand overlaps another one (expr.cu, ring_ok): a copy issued a trip early must not be overtaken by a store.
__device__ __forceinline__ void ring_cp8(unsigned dst, const double* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ring_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ring_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
constexpr int RING_SLOTS = JIT_NRV + JIT_NRA + JIT_NRS;     // values | old adjoints | the per-element seed (vector-valued stages)
#endif

template <int STORE>
__device__ __forceinline__ void stage_body(const StageArgs& a)
{
    constexpr bool BIG = STORE == 2;
    extern __shared__ double s_dyn[];
    __shared__ double s_red[32 * K_ACC];
    __shared__ bool s_last;
#ifdef FADB_JIT
    // programmatic dependent launch (jit.cu launches the specialised kernels with the attribute): the next launch on the
    // stream may become resident now; nothing is read before the previous grid has completed and flushed
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
#ifndef FADB_JIT
    __shared__ LeafDesc s_leaf_sm[BIG ? 1 : MAXL];
    const LeafDesc* s_leaf = BIG ? a.leaves : s_leaf_sm;
    __shared__ DIns s_prog_sm[BIG ? 1 : MAXS];
    const DIns* s_prog = BIG ? a.prog : s_prog_sm;
    if (!BIG) {
        for (int k = threadIdx.x; k < K_NINS; k += blockDim.x) s_prog_sm[k] = a.prog[k];
        for (int k = threadIdx.x; k < K_NLEAVES; k += blockDim.x) s_leaf_sm[k] = a.leaves[k];
        __syncthreads();
    }
#endif

    const bool fwd_out = K_MODE & M_F;     // forward results are stored in this launch
    bool back = K_MODE & M_B;
    double seed_s = a.seed_ptr ? *a.seed_ptr : a.seed;
    bool gated = a.gate && *a.gate != 0.0;
    double sig_s = 1.0;
    if (K_TERM == T_NORMAL && !K_SIG_VEC) {
        sig_s = a.sig_ptr ? *a.sig_ptr : 1.0;
        if (!(sig_s > 0)) gated = true;                       // normal.hpp:147, 231, 344
    }
    if (K_TERM >= T_NORMAL && !a.seed_vec && seed_s == 0.0) back = false;   // normal.hpp:160, cauchy.hpp:151
    double pr_pnz = 1.0, pr_zeros = 0.0, pr_val = 0.0;
    if (K_TERM == T_PROD && back) { pr_pnz = a.red[0]; pr_zeros = a.red[1]; pr_val = a.red[2]; }

    double acc[K_ACC];
#pragma unroll
    for (int c = 0; c < K_ACC; ++c) acc[c] = 0.0;
    double p_nz = 1.0, p_zeros = 0.0;
#ifdef FADB_JIT
    // Broadcast scalars are read ONCE, ahead of the element loop: inside it the compiler must assume the adjoint stores may
    // alias them, reloads them every trip and cannot hoist what depends on them alone (the reciprocal of a scalar divisor:
    // 2 of the 3 reciprocals per element of sum(-0.5 pow<2>((x - c w0) / w1))).  A scalar that this stage assigns keeps
    // its in-program read.
    double sval[K_LSLOTS];
    K_UNROLL
    for (int k = 0; k < K_NLEAVES; ++k) sval[k] = (K_LEAF_SCALAR(k) && !K_LEAF_ASSIGNED(k)) ? K_LVAL(k)[0] : 0.0;
#endif
#if defined(FADB_JIT) && JIT_RING
    __shared__ double s_ring[RING_SLOTS * 128];
    double* const ring = s_ring + threadIdx.x;
    unsigned ring_sa;
    {
        unsigned long long sa64;
        asm("cvta.to.shared.u64 %0, %1;" : "=l"(sa64) : "l"(ring));
        ring_sa = (unsigned)sa64;
    }
#define RINGV(k) ring[JIT_RV[k] * 128]
#define RINGA(k) ring[(JIT_NRV + JIT_RA[k]) * 128]
#define RING_ISSUE_VALS(j)                                                                                  \
    do {                                                                                                    \
        if ((j) < a.N) {                                                                                    \
            K_UNROLL                                                                                        \
            for (int k_ = 0; k_ < K_NLEAVES; ++k_)                                                          \
                if (JIT_RV[k_] >= 0) ring_cp8(ring_sa + JIT_RV[k_] * 1024, K_LVAL(k_) + (j));               \
        }                                                                                                   \
        ring_commit();                                                                                      \
    } while (0)
#define RING_ISSUE_ADJS(j)                                                                                  \
    do {                                                                                                    \
        if ((K_MODE & M_B) && back && (j) < a.N) {                                                          \
            K_UNROLL                                                                                        \
            for (int k_ = 0; k_ < K_NLEAVES; ++k_)                                                          \
                if (JIT_RA[k_] >= 0) ring_cp8(ring_sa + (JIT_NRV + JIT_RA[k_]) * 1024, K_LADJ(k_) + (j));   \
            if (JIT_NRS && a.seed_vec) ring_cp8(ring_sa + (JIT_NRV + JIT_NRA) * 1024, a.seed_vec + (j));               \
        }                                                                                                   \
        ring_commit();                                                                                      \
    } while (0)
#endif

    double v_loc[STORE == 0 ? K_SLOTS : 1], g_loc[STORE == 0 ? K_SLOTS : 1], l_loc[STORE == 0 ? K_LSLOTS : 1];
    const int vs = STORE == 1 ? (int)blockDim.x : 1;             // slot stride
    double* vb = STORE == 0 ? v_loc : STORE == 1 ? s_dyn + threadIdx.x : a.scratch;
    double* gb = STORE == 0 ? g_loc : vb + (size_t)K_NINS * vs;
    double* lb = STORE == 0 ? l_loc : gb + (size_t)K_NINS * vs;
#define V(k) vb[(k) * vs]
#define G(k) gb[(k) * vs]
#define LA(k) lb[(k) * vs]
    const unsigned long long tid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const unsigned long long nth = (unsigned long long)gridDim.x * blockDim.x;
#if defined(FADB_JIT) && JIT_VEC > 1
    // JIT_VEC consecutive elements per thread and trip: every per-element leaf is read with ONE 256-bit
    // load and its adjoint updated with one 256-bit load + store, as in the hand-written kernels
    // (map_reduce.cu, normal.cu).  The planner only asks for this when the program has no placeholder
    // stores; a.vec_ok says whether every per-element leaf pointer is 32-byte aligned.  A trip that is
    // not full (the last elements) or an unaligned operand takes scalar accesses in the same loop.
    static_assert(JIT_VEC == 4, "the vector path moves 256-bit words");
    for (unsigned long long base = tid * JIT_VEC; base < a.N; base += nth * JIT_VEC) {
        const bool full = a.vec_ok && base + JIT_VEC <= a.N;
        double lv[K_LSLOTS][JIT_VEC], la[K_LSLOTS][JIT_VEC];
        K_UNROLL
        for (int k = 0; k < K_NLEAVES; ++k) {
            if (K_LEAF_SCALAR(k)) continue;
            if (full) {
                const double4_t w = K_LEAF_ADJ_MODE(k) == 0 ? ldg256_stream(K_LVAL(k) + base) : ld256(K_LVAL(k) + base);
#pragma unroll
                for (int e = 0; e < JIT_VEC; ++e) lv[k][e] = w.v[e];
                if (K_LEAF_ADJ_MODE(k) == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(K_LADJ(k) + base));
            } else {
#pragma unroll
                for (int e = 0; e < JIT_VEC; ++e) lv[k][e] = base + e < a.N ? K_LVAL(k)[base + e] : 0.0;
            }
            const unsigned long long nb = base + nth * JIT_VEC;       // next trip's lines
            if (nb < a.N) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(K_LVAL(k) + nb));
                if (K_LEAF_ADJ_MODE(k) == 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(K_LADJ(k) + nb));
            }
#pragma unroll
            for (int e = 0; e < JIT_VEC; ++e) la[k][e] = 0.0;
        }
#pragma unroll
        for (int e = 0; e < JIT_VEC; ++e) {
        const unsigned long long i = base + e;
        if (!full && i >= a.N) break;
#else
#if defined(FADB_JIT) && JIT_RING
    RING_ISSUE_VALS(tid);
    RING_ISSUE_ADJS(tid);
#endif
    for (unsigned long long i = tid; i < a.N; i += nth) {
#endif
#if defined(FADB_JIT) && JIT_RING
        ring_wait1();                       // this trip's values have landed (its old adjoints may still be in flight)
#elif defined(FADB_JIT) && JIT_VEC == 1
        {   // One element per trip leaves a thread with too few bytes in flight (ncu on the 10-node chain:
            // long-scoreboard 13.3 of 20 stall cycles per issue, DRAM 37 %).  Start fetching the NEXT
            // trip's operand and adjoint lines now; prefetches hold no registers and merge with the loads.
            const unsigned long long j = i + FADB_JIT_PF_DIST * nth;
            K_UNROLL
            for (int k = 0; k < K_NLEAVES; ++k) {
                if (K_LEAF_SCALAR(k)) continue;
                if (i == tid) {
                    if (K_LEAF_ADJ_MODE(k) == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(K_LADJ(k) + i));
#pragma unroll
                    for (int d = 1; d < FADB_JIT_PF_DIST; ++d)
                        if (i + d * nth < a.N) {
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(K_LVAL(k) + i + d * nth));
                            if (K_LEAF_ADJ_MODE(k) == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(K_LADJ(k) + i + d * nth));
                        }
                }
                if (j < a.N) {
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(K_LVAL(k) + j));
                    if (K_LEAF_ADJ_MODE(k) == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(K_LADJ(k) + j));
                }
            }
        }
#endif
        // ------------------------------------------------ forward (feval)
#if defined(FADB_JIT) && JIT_HAS_CF
        // if_else in an UNROLLED program: the jumps of the interpreter (k = target) become a watermark -- instruction k runs
        // iff k >= f_resume -- so that every instruction keeps its compile-time index (registers, folded switches) and the
        // skipped branch costs one uniform compare per instruction.  Targets only grow along the program, nested or not.
        int f_resume = 0;
#endif
        K_UNROLL
        for (int k = 0; k < K_NINS; ++k) {
#if defined(FADB_JIT) && JIT_HAS_CF
            if (k < f_resume) continue;
#endif
            const DIns I = K_INS(k);
            double r;
            switch (I.code) {
#if defined(FADB_JIT) && JIT_RING
            // I.a >= 0: behind its assignment by instruction I.a (that register); else this thread's ring slot
            case F_LEAF_V: r = I.a >= 0 ? V(I.a >= 0 ? I.a : 0) : (JIT_RV[I.leaf] >= 0 ? RINGV(I.leaf) : K_LOADV(I.leaf)); break;
#else
            case F_LEAF_V: r = K_LOADV(I.leaf); break;              // var_view.hpp:69
#endif
#ifdef FADB_JIT
            case F_LEAF_S: r = K_LEAF_ASSIGNED(I.leaf) ? K_LVAL(I.leaf)[0] : sval[I.leaf]; break;
#else
            case F_LEAF_S: r = K_LVAL(I.leaf)[0]; break;
#endif
            case F_CONST: r = I.imm; break;
            case F_EQ: {                                                   // eq.hpp:89-92
                r = V(I.a);
                if (fwd_out) const_cast<double*>(K_LVAL(I.leaf))[K_LEAF_SCALAR(I.leaf) ? 0 : i] = r;
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
#ifndef FADB_JIT
            // IfElseNode (if_else.hpp:64-68): the condition is a scalar, so the jump is uniform
            case F_JZ: r = 0.0; if (V(I.a) == 0.0) k = I.b - 1; break;
            case F_JMP: r = 0.0; k = I.b - 1; break;
            case F_PHI: r = V(I.leaf) != 0.0 ? V(I.a) : V(I.b); break;
#elif JIT_HAS_CF
            case F_JZ: r = 0.0; if (V(I.a) == 0.0) f_resume = I.b; break;
            case F_JMP: r = 0.0; f_resume = I.b; break;
            case F_PHI: r = V(I.leaf) != 0.0 ? V(I.a) : V(I.b); break;
#endif
            default: {                                                     // OpEqNode::feval, eq.hpp:240-247
                const double old = V(I.b), e = V(I.a);
                const int op = I.code - F_OPEQ;
                r = op == FADB_ADD ? old + e : op == FADB_SUB ? old - e : op == FADB_MUL ? old * e : old / e;
                if (fwd_out) const_cast<double*>(K_LVAL(I.leaf))[i] = r;
                break;
            }
            }
            V(k) = r;
        }
        // ------------------------------------------------ terminal, forward part
        double nrm_d = 0, nrm_s = 1, nrm_z = 0;
        if (K_TERM == T_SUM) acc[0] += V(K_ROOT);
        else if (K_TERM == T_PROD) { if (V(K_ROOT) == 0) p_zeros += 1.0; else p_nz *= V(K_ROOT); }
        else if (K_TERM == T_NORMAL) {
            nrm_d = V(K_NX) - V(K_NM);
            nrm_s = V(K_NS);
            if (K_SIG_VEC) {
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
        } else if (K_TERM == T_CAUCHY) {
            // CauchyAdjLogPDFNode (stat/cauchy.hpp): -log(gamma + d^2/gamma) per element, gamma > 0
            const double d = V(K_NX) - V(K_NM), gm = V(K_NS);
            if (!(gm > 0)) { acc[2] += 1.0; if (!K_SIG_VEC) gated = true; }
            acc[0] += a.N == 1 ? log(gm + (d * d) / gm)                 // cauchy.hpp:144-146
                               : log(gm + (1. / gm) * (d * d));         // cauchy.hpp:216-217
        } else if (K_TERM == T_UNIFORM) {
            // UniformAdjLogPDFNode (stat/uniform.hpp): -log(max - min), min < x < max
            const double x = V(K_NX), lo = V(K_NM), hi = V(K_NS);
            if (!(lo < x && x < hi)) acc[2] += 1.0;
            acc[0] += log(hi - lo);
        } else if (K_TERM == T_BERNOULLI) {
            // BernoulliAdjLogPDFNode (stat/bernoulli.hpp:119-141): x in {0,1}, 0 < p < 1
            const double x = V(K_NX), p = V(K_NM);
            const bool in_range = 0 < p && p < 1, zero_one = x == 0 || x == 1;
            if (!in_range || !zero_one) acc[2] += 1.0;
            if (!in_range) acc[0] += (p <= 0) ? (x == 0 ? 0.0 : -INFINITY) : (x == 1 ? 0.0 : -INFINITY);
            else acc[0] += x == 0 ? log(1 - p) : x == 1 ? log(p) : -INFINITY;
        } else if (fwd_out && a.out_vec) a.out_vec[i] = V(K_ROOT);
#if defined(FADB_JIT) && JIT_RING
        RING_ISSUE_VALS(i + nth);           // the forward sweep holds its operands in registers now
        if (!back) { ring_commit(); continue; }
        ring_wait1();                       // this trip's old adjoints and seed have landed
#endif
        if (!back) continue;

        // ------------------------------------------------ adjoint sweep (beval)
        K_UNROLL
        for (int k = 0; k < K_NINS; ++k) G(k) = 0.0;
        K_UNROLL
        for (int k = 0; k < K_NLEAVES; ++k) LA(k) = 0.0;
#if defined(FADB_JIT) && JIT_RING
        const double sd = a.seed_vec ? (JIT_NRS ? ring[(JIT_NRV + JIT_NRA) * 128] : a.seed_vec[i]) : seed_s;
#else
        const double sd = a.seed_vec ? a.seed_vec[i] : seed_s;
#endif
        if (!gated) {
            if (K_TERM == T_NORMAL) {
                const double x = V(K_NX), m = V(K_NM), s = nrm_s, d = nrm_d;
                if (K_SIG_VEC) {
                    const double s2 = s * s;
                    if (K_MU_VEC) { G(K_NX) += sd * (m - x) / s2; G(K_NM) += sd * d / s2; }      // :563-564
                    else { G(K_NX) += (sd / s2) * (m - x); G(K_NM) += sd * (d / s2); }           // :471-474
                    G(K_NS) += (sd / s) * (nrm_z * nrm_z - 1.);                                  // :468, 560
                } else if (a.N == 1) {
                    const double inv_s = 1. / s, z = d * inv_s;                                  // :158-171
                    G(K_NS) += sd * (z * z - 1) * inv_s;
                    const double adj = sd * z * inv_s;
                    G(K_NM) += adj;
                    G(K_NX) += -adj;
                } else {
                    const double inv_s = 1. / s, inv_s_sq = inv_s * inv_s, c = sd * inv_s_sq;
                    G(K_NX) += c * (m - x);                                                      // :282, 371
                    G(K_NM) += K_MU_VEC ? c * d : sd * (d * inv_s_sq);                           // :370, 277
                    G(K_NS) += sd * ((d * d) / (s * s) - 1.) * inv_s;                            // :273 per element
                }
            } else if (K_TERM == T_CAUCHY) {
                const double d = V(K_NX) - V(K_NM), gm = V(K_NS);
                if (a.N == 1) {                                                                   // cauchy.hpp:149-165
                    const double inner = gm + (d * d) / gm;
                    const double x0_adj = 2. * d / (gm * inner);
                    G(K_NS) += sd * (1. / gm * (x0_adj * d - 1));
                    G(K_NM) += sd * x0_adj;
                    G(K_NX) += sd * -x0_adj;
                } else {   // vector cases exactly as written in cauchy.hpp:220-237 (seed enters dx0, dgamma twice)
                    const double dx = (-2. * sd) * d / (gm * gm + d * d);
                    G(K_NS) += (-sd / gm) * (dx * d + 1);
                    G(K_NM) += (-sd) * dx;
                    G(K_NX) += dx;
                }
            } else if (K_TERM == T_UNIFORM) {                                                     // uniform.hpp:157-163
                const double adj = sd / (V(K_NS) - V(K_NM));
                G(K_NS) += -adj;
                G(K_NM) += adj;
            } else if (K_TERM == T_BERNOULLI) {                                                   // bernoulli.hpp:143-149
                const double x = V(K_NX), p = V(K_NM);
                G(K_NM) += x == 0 ? -sd / (1 - p) : sd / p;
            } else if (K_TERM == T_PROD) {
                const double e = V(K_ROOT);                                                      // prod.hpp:194-207
                G(K_ROOT) = (e == 0) ? sd * ((pr_zeros > 1) ? 0.0 : pr_pnz) : sd * (pr_val / e);
            } else {
                G(K_ROOT) = sd;
            }
#if defined(FADB_JIT) && JIT_HAS_CF
            int b_lim = K_NINS;          // backward watermark: instruction k is swept iff k <= b_lim (IfElseNode::beval, if_else.hpp:70-78)
#endif
            K_UNROLL
            for (int k = K_NINS - 1; k >= 0; --k) {
#if defined(FADB_JIT) && JIT_HAS_CF
                if (k > b_lim) continue;
#endif
                const DIns I = K_INS(k);
                const double s = G(k);
                switch (I.code) {
                case F_LEAF_V: case F_LEAF_S: LA(I.leaf) += s; break;
                case F_CONST: break;
                case F_EQ: {                                                                      // eq.hpp:101-107
                    if (K_LEAF_SCALAR(I.leaf) && a.N > 1) { G(I.a) += s + LA(I.leaf); break; }   // not produced by the planner
#if defined(FADB_JIT) && JIT_RING
                    const double tot = ((JIT_RA[I.leaf] >= 0 ? RINGA(I.leaf) : K_LADJ(I.leaf)[i]) + LA(I.leaf)) + s;
                    if (JIT_RA[I.leaf] >= 0) RINGA(I.leaf) = tot;       // a second assignment of the same placeholder reads this
#else
                    const double tot = (K_LADJ(I.leaf)[i] + LA(I.leaf)) + s;
#endif
                    K_LADJ(I.leaf)[i] = tot;
                    LA(I.leaf) = 0.0;
                    G(I.a) += tot;
                    break;
                }
                case F_GLUE: G(I.b) += s; break;                                                  // lhs gets 0, glue.hpp:81-86
                case F_POW: G(I.a) += pow_bmap((long)I.leaf, s, V(I.a), V(k)); break;            // pow.hpp:101-162
                case F_POW2: { const double x = V(I.a); G(I.a) += 2. * s * (x == 0 ? 0 : div_guarded(V(k), x)); break; }
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
                case F_BINARY + FADB_DIV: { double qr, ql; div2_guarded(-s * V(k), s, V(I.b), qr, ql); G(I.b) += qr; G(I.a) += ql; break; }
                case F_BINARY + 12: G(I.b) += -2. * s; G(I.a) += s; break;                        // MockBinary
                // comparisons: no adjoint (binary.hpp:124)
                case F_BINARY + FADB_LT: case F_BINARY + FADB_LE: case F_BINARY + FADB_GT: case F_BINARY + FADB_GE:
                case F_BINARY + FADB_EQ: case F_BINARY + FADB_NE: case F_BINARY + FADB_AND: case F_BINARY + FADB_OR: break;
#ifndef FADB_JIT
                // IfElseNode::beval (if_else.hpp:70-78): only the taken branch is swept
                case F_PHI:
                    if (V(I.leaf) != 0.0) { G(I.a) += s; k = (int)I.imm + 1; }    // skip the else branch
                    else G(I.b) += s;
                    break;
                case F_JMP: if (V(I.a) == 0.0) k = I.leaf + 1; break;             // else was taken: skip the if branch
                case F_JZ: break;
#elif JIT_HAS_CF
                case F_PHI:
                    if (V(I.leaf) != 0.0) { G(I.a) += s; b_lim = (int)I.imm; }    // skip the else branch: resume at its JMP
                    else G(I.b) += s;
                    break;
                case F_JMP: if (V(I.a) == 0.0) b_lim = I.leaf; break;             // else was taken: skip the if branch, resume at its JZ
                case F_JZ: break;
#endif
                default: {                                                         // OpEqNode::beval, eq.hpp:249-271
                    const double tot = (K_LADJ(I.leaf)[i] + LA(I.leaf)) + s;
                    LA(I.leaf) = 0.0;
                    const double old = V(I.b), e = V(I.a);
                    const int op = I.code - F_OPEQ;
                    double ls, rs;
                    if (op == FADB_ADD) { ls = tot; rs = tot; }
                    else if (op == FADB_SUB) { ls = tot; rs = -tot; }
                    else if (op == FADB_MUL) { ls = tot * e; rs = tot * old; }
                    else { ls = tot / e; rs = -tot * old / (e * e); }
                    G(I.a) += rs;
                    K_LADJ(I.leaf)[i] = ls;                                                 // reset_adj(); beval(lseed)
                    const_cast<double*>(K_LVAL(I.leaf))[i] = old;                           // the previous value is restored
                    break;
                }
                }
            }
        }
        // ------------------------------------------------ leaf write-back: adj += seed (var_view.hpp:91-94)
        K_UNROLL
        for (int k = 0; k < K_NLEAVES; ++k) {
            if (K_LEAF_ADJ_MODE(k) == 0) continue;
            if (K_LEAF_ASSIGNED(k)) {  // reads that precede the assignment in program order
                if (LA(k) != 0.0) K_LADJ(k)[i] += LA(k);
                continue;
            }
            if (K_LEAF_SCALAR(k) && a.N > 1) { acc[K_LEAF_CHANNEL(k) < 0 ? 0 : K_LEAF_CHANNEL(k)] += LA(k); continue; }
#if defined(FADB_JIT) && JIT_VEC > 1
            la[k][e] = LA(k);                              // stored by the vector write-back below
#else
            if (K_LEAF_ADJ_MODE(k) == 2) K_LADJ(k)[i] = LA(k);
#if defined(FADB_JIT) && JIT_RING
            else if (!gated) K_LADJ(k)[i] = (JIT_RA[k] >= 0 ? RINGA(k) : K_LADJ(k)[i]) + LA(k);
#else
            else if (!gated) K_LADJ(k)[i] += LA(k);
#endif
#endif
        }
#if defined(FADB_JIT) && JIT_RING
        RING_ISSUE_ADJS(i + nth);
#endif
#if defined(FADB_JIT) && JIT_VEC > 1
        }   // e
        if (back) {
            K_UNROLL
            for (int k = 0; k < K_NLEAVES; ++k) {
                if (K_LEAF_SCALAR(k) || K_LEAF_ADJ_MODE(k) == 0) continue;
                if (K_LEAF_ADJ_MODE(k) == 1 && gated) continue;
                if (full) {
                    double4_t w;
                    if (K_LEAF_ADJ_MODE(k) == 1) {
                        w = ld256(K_LADJ(k) + base);
#pragma unroll
                        for (int e = 0; e < JIT_VEC; ++e) w.v[e] += la[k][e];
                    } else {
#pragma unroll
                        for (int e = 0; e < JIT_VEC; ++e) w.v[e] = la[k][e];
                    }
                    st256(K_LADJ(k) + base, w);
                } else {
#pragma unroll
                    for (int e = 0; e < JIT_VEC; ++e)
                        if (base + e < a.N) {
                            if (K_LEAF_ADJ_MODE(k) == 2) K_LADJ(k)[base + e] = la[k][e];
                            else K_LADJ(k)[base + e] += la[k][e];
                        }
                }
            }
        }
#endif
    }

    // ---------------------------------------------------- grid reduction + scalar epilogue
    if (K_NCH == 0 && K_TERM != T_PROD) return;
    if (K_TERM == T_PROD && (K_MODE & M_F)) {
        const double bp = block_prod_all(p_nz, s_red);
        acc[0] = p_zeros;
        block_sum<K_ACC>(acc, s_red);
        if (threadIdx.x == 0) {
            a.partials[(size_t)blockIdx.x * (K_ACC + 1) + K_ACC] = bp;
            for (int c = 0; c < K_ACC; ++c) a.partials[(size_t)blockIdx.x * (K_ACC + 1) + c] = acc[c];
            __threadfence();
            s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        double p = 1.0, z = 0.0;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
            p *= __ldcg(&a.partials[(size_t)b * (K_ACC + 1) + K_ACC]);
            z += __ldcg(&a.partials[(size_t)b * (K_ACC + 1)]);
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
    block_sum<K_ACC>(acc, s_red);
    if (threadIdx.x == 0) {
        for (int c = 0; c < K_ACC; ++c) a.partials[(size_t)blockIdx.x * (K_ACC + 1) + c] = acc[c];
        __threadfence();
        s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double tot[K_ACC];
#pragma unroll
    for (int c = 0; c < K_ACC; ++c) tot[c] = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x)
#pragma unroll
        for (int c = 0; c < K_ACC; ++c) tot[c] += __ldcg(&a.partials[(size_t)b * (K_ACC + 1) + c]);
    block_sum<K_ACC>(tot, s_red);
    if (threadIdx.x != 0) return;
    *a.ticket = 0u;
    if (K_MODE & M_F) {
        if (K_TERM == T_SUM) a.out_scalar[0] = tot[0];
        else if (K_TERM == T_NORMAL) {
            double val;
            if (K_SIG_VEC) {
                a.red[1] = tot[2];
                val = (tot[2] != 0.0) ? -INFINITY : -0.5 * tot[0] - tot[1];                      // normal.hpp:451, 548
            } else {
                a.red[1] = (sig_s > 0) ? 0.0 : 1.0;
                if (!(sig_s > 0)) val = -INFINITY;
                else if (a.N == 1) val = -0.5 * tot[0] - log(sig_s);
                else val = -0.5 * (tot[0] / (sig_s * sig_s)) - (double)a.N * log(sig_s);         // normal.hpp:245-246
            }
            a.out_scalar[0] = val;
        } else if (K_TERM == T_CAUCHY || K_TERM == T_UNIFORM) {
            a.red[1] = tot[2];
            a.out_scalar[0] = tot[2] != 0.0 ? -INFINITY : -tot[0];
        } else if (K_TERM == T_BERNOULLI) {
            a.red[1] = tot[2];
            a.out_scalar[0] = tot[0];
        }
    }
    if (K_MODE & M_B) {
        K_UNROLL
        for (int k = 0; k < K_NLEAVES; ++k) {
            if (K_LEAF_CHANNEL(k) < 0 || K_LEAF_ADJ_MODE(k) == 0) continue;
            const bool skip = gated || !back || (K_TERM > T_NORMAL && (K_MODE & M_F) && tot[2] != 0.0);
            if (K_LEAF_ADJ_MODE(k) == 2) K_LADJ(k)[0] = skip ? 0.0 : tot[K_LEAF_CHANNEL(k)];
            else if (!skip) K_LADJ(k)[0] += tot[K_LEAF_CHANNEL(k)];
        }
    }
}

#undef V
#undef G
#undef LA

#ifdef FADB_JIT
#ifndef FADB_JIT_MINB
#define FADB_JIT_MINB 8   // <= 64 registers: measured best on B200 (chain 0.127 -> 0.113 ms, Black-Scholes stage 0.070 -> 0.061 ms vs no cap)
#endif
extern "C" __global__ void __launch_bounds__(128, FADB_JIT_MINB) fadb_stage_jit(StageArgs a) { stage_body<0>(a); }
#else
template <int STORE>
__global__ void __launch_bounds__(128) stage_kernel(StageArgs a) { stage_body<STORE>(a); }
#endif

}  // namespace fadb
