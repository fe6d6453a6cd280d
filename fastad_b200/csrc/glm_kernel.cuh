// glm_kernel.cuh — ONE pass over X for  sum(f(dot(X, w), row operands...))  and  normal_adj_log_pdf(y, f(dot(X, w), ...), sigma)
// with one sigma  (NVRTC only, csrc/jit.cu kind 1).
//
// Replaces the three launches DotNode::feval (dot.hpp:89-95, X streamed once) -> the elementwise stage with its
// SumElemNode (sum.hpp:156-173) -> DotNode::beval (dot.hpp:96-104, X streamed AGAIN) by one sweep: with the root
// seed known up front the adjoint of row i's product (X w)_i depends on that row alone, so a 64-row tile that has
// just produced its 64 products can evaluate the stage program forward AND backward for those rows and
// accumulate X^T a from the same registers -- X crosses HBM once (the shape of a GLM log-likelihood:
// regression with a link function).  linreg.cu's TMA kernel is the fixed-function instance of this (f = normal
// log-density / squared residual); here f is the planner's stage program, spliced in as compile-time constants
// exactly as for stage_kernel.cuh (JIT_PROG, JIT_LEAF ...), plus
//   GLM_DOT_LEAF            leaf index of the dot result inside the program,
//   GLM_NVEC, GLM_VEC_SLOT  per-row constant operands (y ...): staged through the ring next to the X tile,
//   scalar operands         read once; the adjoints of scalar VARIABLES are reduction channels 1.. (JIT_LEAF.channel).
// Layout and pipeline are linreg_tma_kernel's: one producer warp issues ONE cp.async.bulk.tensor.2d per 64 x cols
// tile (+ one bulk copy per row operand) into a 3-stage mbarrier ring; four consumer warps own 16 columns each,
// lane l owns rows {2l, 2l+1}.  Every consumer warp evaluates f for the tile's 64 rows (each needs the row
// adjoints for its own columns; the redundancy costs FP64 issue slots the memory-bound sweep does not use) -- unless the map is
// expensive (GLM_SPLIT, chosen by the host from the program's transcendental count): then warp q evaluates rows 16q .. 16q+15 only,
// one row per lane, and the row adjoints travel through shared memory (one more barrier per tile, a quarter of the map work).
#ifdef FADB_JIT
namespace fadb {

struct alignas(64) GlmTensorMap { unsigned long long opaque[16]; };   // CUtensorMap (cuda.h) without the header

constexpr int GL_ROWS = 64, GL_STAGES = 3, GL_CONSUMERS = 128, GL_THREADS = GL_CONSUMERS + 32, GL_CPW = 16;
constexpr int GL_NV = GLM_NVEC > 0 ? GLM_NVEC : 1;
constexpr int GL_STAGE_BYTES = 64 * GL_ROWS * 8 + GL_NV * GL_ROWS * 8;
constexpr int GL_ACC = JIT_NCH > 1 ? JIT_NCH : 1;     // channel 0: sum of the row values; 1..: adjoints of scalar variables
constexpr int GL_NL = JIT_NLEAVES > 0 ? JIT_NLEAVES : 1;
#ifndef GLM_SPLIT
#define GLM_SPLIT 0
#endif

struct GlmArgs {
    unsigned long long rows;
    int cols, pad;
    const double* w;
    double seed;
    double* cta_partials;      // [grid][GL_ACC + cols]
    unsigned int* ticket;
    double* out;               // [GL_ACC + cols]: channel totals, then X^T a
    const double* lp_val[MAXL_JIT];
};

__device__ __forceinline__ void gl_consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(GL_CONSUMERS) : "memory"); }
__device__ __forceinline__ void gl_tma_load_2d(void* dst_smem, const GlmTensorMap* map, int c0, int c1, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst_smem)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

// The stage program for ONE row: forward, then the adjoint sweep from the root seed.  t = the row's product (X w)_i,
// lv = its row operands, sv = the scalar operands.  Same instruction semantics as stage_body (stage_kernel.cuh).
// the ONE sigma of a normal terminal and what the row formulas need of it, formed once per thread (IEEE divisions)
struct GlmSigma { double inv_s, inv_s_sq, inv_ss; bool valid; };

__device__ __forceinline__ void glm_row(double t, const double (&lv)[GL_NV], const double (&sv)[GL_NL], double seed, const GlmSigma& sig,
                                        bool tally, double (&acc)[GL_ACC], double& adj)
{
    double v[JIT_NINS], g[JIT_NINS], la[GL_NL];
#pragma unroll
    for (int k = 0; k < JIT_NINS; ++k) {
        const DIns I = JIT_PROG[k];
        double r;
        switch (I.code) {
        case F_LEAF_V: r = I.leaf == GLM_DOT_LEAF ? t : lv[GLM_VEC_SLOT[I.leaf] < 0 ? 0 : GLM_VEC_SLOT[I.leaf]]; break;
        case F_LEAF_S: r = sv[I.leaf]; break;
        case F_CONST: r = I.imm; break;
        case F_POW: r = pow_int((long)I.leaf, v[I.a]); break;
        case F_POW2: r = v[I.a] * v[I.a]; break;
#define U(FN) case F_UNARY + FN: r = Unary<FN>::f(v[I.a]); break;
        U(FADB_NEG) U(FADB_SIN) U(FADB_COS) U(FADB_TAN) U(FADB_ASIN) U(FADB_ACOS) U(FADB_ATAN) U(FADB_EXP)
        U(FADB_LOG) U(FADB_SQRT) U(FADB_ERF) U(FADB_SIGMOID) U(FADB_SINH) U(FADB_COSH) U(FADB_TANH) U(FADB_SQUARE)
#undef U
        case F_UNARY + 16: r = 2 * v[I.a]; break;
        case F_BINARY + FADB_ADD: r = v[I.a] + v[I.b]; break;
        case F_BINARY + FADB_SUB: r = v[I.a] - v[I.b]; break;
        case F_BINARY + FADB_MUL: r = v[I.a] * v[I.b]; break;
        case F_BINARY + FADB_DIV: r = div_guarded(v[I.a], v[I.b]); break;
        case F_BINARY + FADB_LT: r = v[I.a] < v[I.b] ? 1.0 : 0.0; break;
        case F_BINARY + FADB_LE: r = v[I.a] <= v[I.b] ? 1.0 : 0.0; break;
        case F_BINARY + FADB_GT: r = v[I.a] > v[I.b] ? 1.0 : 0.0; break;
        case F_BINARY + FADB_GE: r = v[I.a] >= v[I.b] ? 1.0 : 0.0; break;
        case F_BINARY + FADB_EQ: r = v[I.a] == v[I.b] ? 1.0 : 0.0; break;
        case F_BINARY + FADB_NE: r = v[I.a] != v[I.b] ? 1.0 : 0.0; break;
        case F_BINARY + FADB_AND: r = (v[I.a] != 0 && v[I.b] != 0) ? 1.0 : 0.0; break;
        case F_BINARY + FADB_OR: r = (v[I.a] != 0 || v[I.b] != 0) ? 1.0 : 0.0; break;
        case F_BINARY + 12: r = v[I.a] - 2 * v[I.b]; break;
        default: r = NAN; break;                       // the planner routes nothing else here
        }
        v[k] = r;
        g[k] = 0.0;
    }
#pragma unroll
    for (int k = 0; k < GL_NL; ++k) la[k] = 0.0;
    double term_val;
    bool valid = true;
    if (JIT_TERM == T_NORMAL) {
        // NormalAdjLogPDFNode with a per-row mean and ONE sigma (normal.hpp:304-381), formulas as in stage_body's terminal
        const double x = v[JIT_NX], m = v[JIT_NM], d = x - m;
        term_val = d * d;                              // squaredNorm, normal.hpp:351; the epilogue forms the value
        valid = sig.valid;                             // normal.hpp:344: sigma <= 0 => -inf and no adjoint anywhere
        const double c = seed * sig.inv_s_sq;
        g[JIT_NX] += c * (m - x);                      // normal.hpp:371
        g[JIT_NM] += c * d;                            // normal.hpp:370
        // normal.hpp:365 per element, (d d) / (s s) as a product with the reciprocal formed once: the term only enters the
        // REDUCED sigma adjoint (1e-12 class), the per-row mean adjoint above is formed exactly as the stage kernel forms it
        g[JIT_NS] += seed * ((d * d) * sig.inv_ss - 1.) * sig.inv_s;
    } else {
        constexpr int root = JIT_ROOT < 0 ? 0 : JIT_ROOT;
        term_val = v[root];
        g[root] = seed;                                // SumElemNode::beval: every element gets the seed (sum.hpp:169-172)
    }
#pragma unroll
    for (int k = JIT_NINS - 1; k >= 0; --k) {
        const DIns I = JIT_PROG[k];
        const double s = g[k];
        switch (I.code) {
        case F_LEAF_V: case F_LEAF_S: la[I.leaf] += s; break;
        case F_POW: g[I.a] += pow_bmap((long)I.leaf, s, v[I.a], v[k]); break;
        case F_POW2: { const double x = v[I.a]; g[I.a] += 2. * s * (x == 0 ? 0 : div_guarded(v[k], x)); break; }
#define U(FN) case F_UNARY + FN: g[I.a] += Unary<FN>::b(s, v[I.a], v[k]); break;
        U(FADB_NEG) U(FADB_SIN) U(FADB_COS) U(FADB_TAN) U(FADB_ASIN) U(FADB_ACOS) U(FADB_ATAN) U(FADB_EXP)
        U(FADB_LOG) U(FADB_SQRT) U(FADB_ERF) U(FADB_SIGMOID) U(FADB_SINH) U(FADB_COSH) U(FADB_TANH) U(FADB_SQUARE)
#undef U
        case F_UNARY + 16: g[I.a] += s * 2; break;
        case F_BINARY + FADB_ADD: g[I.b] += s; g[I.a] += s; break;
        case F_BINARY + FADB_SUB: g[I.b] += -s; g[I.a] += s; break;
        case F_BINARY + FADB_MUL: { const double x = v[I.a], y = v[I.b]; g[I.b] += s * x; g[I.a] += s * y; break; }
        case F_BINARY + FADB_DIV: { double qr, ql; div2_guarded(-s * v[k], s, v[I.b], qr, ql); g[I.b] += qr; g[I.a] += ql; break; }
        case F_BINARY + 12: g[I.b] += -2. * s; g[I.a] += s; break;
        default: break;                                // constants, comparisons: no adjoint
        }
    }
    adj = valid ? la[GLM_DOT_LEAF] : 0.0;
    if (tally) {
        acc[0] += term_val;
#pragma unroll
        for (int k = 0; k < JIT_NLEAVES; ++k)
            if (JIT_LEAF[k].scalar && JIT_LEAF[k].adj_mode != 0 && JIT_LEAF[k].channel > 0) acc[JIT_LEAF[k].channel] += valid ? la[k] : 0.0;
    }
}

extern "C" __global__ void __launch_bounds__(GL_THREADS, 2)
fadb_glm_jit(GlmArgs a, const __grid_constant__ GlmTensorMap xmap)
{
    extern __shared__ __align__(128) unsigned char gl_smem[];
    unsigned char* stages = gl_smem;                                              // [GL_STAGES][GL_STAGE_BYTES]
    double* w_s = reinterpret_cast<double*>(gl_smem + GL_STAGES * GL_STAGE_BYTES); // [64]
    double* pt_s = w_s + 64;                                                      // [2][4][GL_ROWS]
    double* g_s = pt_s + 2 * 4 * GL_ROWS;                                         // [64]
    double* ch_s = g_s + 64;                                                      // [4][16] channel totals per consumer warp
    double* a_s = ch_s + 64;                                                      // [GL_ROWS] row adjoints of the tile (GLM_SPLIT)
    uint64_t* full = reinterpret_cast<uint64_t*>(a_s + GL_ROWS);                  // [GL_STAGES]
    uint64_t* empty = full + GL_STAGES;                                           // [GL_STAGES]
    __shared__ bool is_last;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cols = a.cols;
    asm volatile("griddepcontrol.launch_dependents;");        // programmatic dependent launch, as the library's own kernels
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x < 64) w_s[threadIdx.x] = (int)threadIdx.x < cols ? a.w[threadIdx.x] : 0.0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < GL_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 4); }
        fence_mbar_init();
    }
    __syncthreads();

    const unsigned long long ntiles = (a.rows + GL_ROWS - 1) / GL_ROWS;
    if (warp == 4) {
        // ------------------------------------------------ producer warp
        int stage = 0;
        unsigned phase = 0;
        for (unsigned long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            mbar_wait(&empty[stage], phase ^ 1);
            const unsigned long long r0 = tile * GL_ROWS;
            const unsigned nrow = (unsigned)((a.rows - r0 < (unsigned long long)GL_ROWS) ? a.rows - r0 : GL_ROWS);
            unsigned char* base = stages + (size_t)stage * GL_STAGE_BYTES;
            if (lane == 0) {
                mbar_arrive_expect_tx(&full[stage], (unsigned)cols * GL_ROWS * 8u + (unsigned)GLM_NVEC * nrow * 8u);
                gl_tma_load_2d(base, &xmap, (int)r0, 0, &full[stage]);      // rows beyond the matrix are zero-filled
#pragma unroll
                for (int k = 0; k < JIT_NLEAVES; ++k)
                    if (GLM_VEC_SLOT[k] >= 0)
                        bulk_g2s(base + 64 * GL_ROWS * 8 + GLM_VEC_SLOT[k] * GL_ROWS * 8, a.lp_val[k] + r0, nrow * 8u, &full[stage]);
            }
            if (++stage == GL_STAGES) { stage = 0; phase ^= 1; }
        }
        return;
    }
    // ---------------------------------------------------- consumer warps
    double sv[GL_NL];
#pragma unroll
    for (int k = 0; k < GL_NL; ++k) sv[k] = (k < JIT_NLEAVES && JIT_LEAF[k].scalar) ? a.lp_val[k][0] : 0.0;
    GlmSigma sig{1.0, 1.0, 1.0, true};
    if (JIT_TERM == T_NORMAL) {
        constexpr int ns = JIT_NS < 0 ? 0 : JIT_NS;
        const double sg = JIT_PROG[ns].code == F_CONST ? JIT_PROG[ns].imm : sv[JIT_PROG[ns].leaf < 0 ? 0 : JIT_PROG[ns].leaf];
        sig.valid = sg > 0;
        sig.inv_s = 1. / sg;                            // normal.hpp:354-355
        sig.inv_s_sq = sig.inv_s * sig.inv_s;
        sig.inv_ss = 1. / (sg * sg);
    }
    double g[GL_CPW];
#pragma unroll
    for (int c = 0; c < GL_CPW; ++c) g[c] = 0.0;
    double acc[GL_ACC];
#pragma unroll
    for (int c = 0; c < GL_ACC; ++c) acc[c] = 0.0;
    int stage = 0;
    unsigned phase = 0, par = 0;
    for (unsigned long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const unsigned long long r0 = tile * GL_ROWS + 2 * lane;
        const bool in0 = r0 < a.rows, in1 = r0 + 1 < a.rows;
        mbar_wait(&full[stage], phase);
        const unsigned char* base = stages + (size_t)stage * GL_STAGE_BYTES;
        double2 xv[GL_CPW];
#pragma unroll
        for (int c = 0; c < GL_CPW; ++c) {
            const int j = warp * GL_CPW + c;
            double2 x2 = make_double2(0.0, 0.0);
            if (j < cols) x2 = *reinterpret_cast<const double2*>(base + (size_t)j * GL_ROWS * 8 + lane * 16);
            if (!in0) x2.x = 0.0;
            if (!in1) x2.y = 0.0;
            xv[c] = x2;
        }
        double lv0[GL_NV], lv1[GL_NV];
#pragma unroll
        for (int q = 0; q < GL_NV; ++q) {
            double2 y2 = make_double2(1.0, 1.0);
            if (q < GLM_NVEC) y2 = *reinterpret_cast<const double2*>(base + 64 * GL_ROWS * 8 + q * GL_ROWS * 8 + lane * 16);
            lv0[q] = in0 ? y2.x : 1.0;
            lv1[q] = in1 ? y2.y : 1.0;
        }
#if GLM_SPLIT
        const int row_e = 16 * warp + (lane & 15);     // the row this lane evaluates (lanes 16..31 shadow lanes 0..15)
        const bool in_e = tile * GL_ROWS + row_e < a.rows;
        double lve[GL_NV];
#pragma unroll
        for (int q = 0; q < GL_NV; ++q) {
            double yv = 1.0;
            if (q < GLM_NVEC) yv = *reinterpret_cast<const double*>(base + 64 * GL_ROWS * 8 + q * GL_ROWS * 8 + row_e * 8);
            lve[q] = in_e ? yv : 1.0;
        }
#endif
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);     // the tile is in registers: free the slot
        if (++stage == GL_STAGES) { stage = 0; phase ^= 1; }

        double s0a = 0.0, s0b = 0.0, s1a = 0.0, s1b = 0.0;
#pragma unroll
        for (int c = 0; c < GL_CPW; c += 2) {
            const double wa = w_s[warp * GL_CPW + c], wb = w_s[warp * GL_CPW + c + 1];
            s0a += xv[c].x * wa; s0b += xv[c + 1].x * wb;
            s1a += xv[c].y * wa; s1b += xv[c + 1].y * wb;
        }
        double* pt = pt_s + par * 4 * GL_ROWS;
        *reinterpret_cast<double2*>(pt + warp * GL_ROWS + 2 * lane) = make_double2(s0a + s0b, s1a + s1b);
        gl_consumer_bar();
        double a0, a1;
#if GLM_SPLIT
        {
            const double te = (pt[row_e] + pt[GL_ROWS + row_e]) + (pt[2 * GL_ROWS + row_e] + pt[3 * GL_ROWS + row_e]);
            double ae;
            glm_row(te, lve, sv, a.seed, sig, lane < 16 && in_e, acc, ae);
            if (lane < 16) a_s[row_e] = in_e ? ae : 0.0;
        }
        par ^= 1;
        gl_consumer_bar();                             // the tile's 64 row adjoints are in shared memory
        {
            const double2 a2 = *reinterpret_cast<const double2*>(a_s + 2 * lane);
            a0 = a2.x; a1 = a2.y;
        }
#else
        const double2 p0 = *reinterpret_cast<const double2*>(pt + 2 * lane);
        const double2 p1 = *reinterpret_cast<const double2*>(pt + GL_ROWS + 2 * lane);
        const double2 p2 = *reinterpret_cast<const double2*>(pt + 2 * GL_ROWS + 2 * lane);
        const double2 p3 = *reinterpret_cast<const double2*>(pt + 3 * GL_ROWS + 2 * lane);
        par ^= 1;
        glm_row((p0.x + p1.x) + (p2.x + p3.x), lv0, sv, a.seed, sig, warp == 0 && in0, acc, a0);
        glm_row((p0.y + p1.y) + (p2.y + p3.y), lv1, sv, a.seed, sig, warp == 0 && in1, acc, a1);
#endif
        if (!in0) a0 = 0.0;
        if (!in1) a1 = 0.0;
#pragma unroll
        for (int c = 0; c < GL_CPW; ++c) g[c] += xv[c].x * a0 + xv[c].y * a1;
    }
    // ---- CTA reduce: lanes -> column totals and channel totals (consumer threads only from here on)
#pragma unroll
    for (int c = 0; c < GL_CPW; ++c) {
        const double t = warp_sum(g[c]);
        if (lane == 0) g_s[warp * GL_CPW + c] = t;
    }
#pragma unroll
    for (int c = 0; c < GL_ACC; ++c) {                 // every consumer warp tallied rows (GLM_SPLIT) or warp 0 alone did
        const double t = warp_sum(acc[c]);
        if (lane == 0) ch_s[warp * 16 + c] = t;
    }
    gl_consumer_bar();
    const int nslot = GL_ACC + cols;
    double* mine = a.cta_partials + (size_t)blockIdx.x * nslot;
    if ((int)threadIdx.x < nslot)
        mine[threadIdx.x] = (int)threadIdx.x < GL_ACC
                                ? (ch_s[threadIdx.x] + ch_s[16 + threadIdx.x]) + (ch_s[32 + threadIdx.x] + ch_s[48 + threadIdx.x])
                                : g_s[threadIdx.x - GL_ACC];
    gl_consumer_bar();
    if (threadIdx.x == 0) {       // ONE thread publishes the row: its fence is cumulative over the stores the barrier ordered before it
        __threadfence();
        is_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
        if (is_last) __threadfence();
    }
    gl_consumer_bar();
    if (!is_last) return;
    // last CTA: warp q sums rows q, q + 4, ... (lane = slot, slot + 32, slot + 64), then the four warp sums in warp order
    double acc3[3] = {0.0, 0.0, 0.0};
    for (unsigned b = warp; b < gridDim.x; b += 4)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int slot = lane + 32 * k;
            if (slot < nslot) acc3[k] += __ldcg(&a.cta_partials[(size_t)b * nslot + slot]);
        }
    double* red = pt_s;                                   // [4][96] of its 512 doubles (the tile loop is over)
#pragma unroll
    for (int k = 0; k < 3; ++k) red[warp * 96 + lane + 32 * k] = acc3[k];
    gl_consumer_bar();
    if ((int)threadIdx.x < nslot)
        a.out[threadIdx.x] = ((red[threadIdx.x] + red[96 + threadIdx.x]) + red[192 + threadIdx.x]) + red[288 + threadIdx.x];
    if (threadIdx.x == 0) *a.ticket = 0u;
}

}  // namespace fadb
#endif
