// mlp3.cu — K6: 3-layer network with jump connection, batched forward + reverse sweep.
//
// Replaces the per-row loop `for each row: copy row; autodiff(expr, seed)` of
// example/ceres_3layer_neural_net/src.cpp:48-55 over the expression of README.md:706-711
// (src.cpp:38-40):   pow<2>(yi - (dot(A3, sigmoid(dot(A2, sigmoid(x1)) + b2) + x1) + b3)),
// x1 = dot(A1, xi) + b1, i.e. DotNode (dot.hpp:89-104), UnaryNode<Sigmoid> (unary.hpp:273-278),
// BinaryNode<Add/Sub> (binary.hpp:266-303), PowNode<2> (pow.hpp:91-140) feval + beval, with the
// 25 parameter adjoints accumulated over rows by VarView::beval (var_view.hpp:91-94).
// One thread owns one row: forward and adjoint sweep in registers; the 25 gradients + loss are
// reduced across the CTA with warp shuffles and across CTAs by the last-arriving CTA.
// The matrices are 3x2, 3x3 and 1x3: far below a DMMA tile, so the FP64 tensor path is not used.
#include <algorithm>
#include <type_traits>

#define FM_CONST_LITERALS      // fastmath.cuh: exp_t's literals from the constant bank (measured: 100.8 -> 97.5 us at 2^22 rows)
#include "common.cuh"
#include "p2p_device.cuh"
#include "../../include/fastad_b200/fastmath.cuh"

namespace fadb {

constexpr int NP = 25;            // parameters: A1(3x2) b1(3) A2(3x3) b2(3) A3(1x3) b3(1), column-major
constexpr int NOUT = NP + 1;      // + loss

// exp on the ordinary range through fastmath.cuh (<= 1 ulp), libdevice otherwise
// 1 / (1 + e) by Newton reciprocal; e = +inf (x < -709) gives 0 like the IEEE division would
__device__ __forceinline__ double inv1p(double e) { return e > 1e300 ? 0.0 : fm::rcp_(1 + e); }
// t32 != nullptr: the 32-entry-table exp of fastmath.cuh (11 instead of 20 FP64 instructions, <= 1 ulp)
__device__ __forceinline__ double exp_s(double x, const double* t32)
{
    return (fabs(x) <= 700.0) ? (t32 ? fm::exp_t(x, t32) : fm::exp_(x)) : exp(x);
}

// one row: forward + adjoint sweep in registers, the 26 sums accumulated into acc
template <bool TAB>
__device__ __forceinline__ void mlp3_row(const double* __restrict__ p, const double* __restrict__ t32s, double xa, double xb, double yy,
                                         double (&acc)[NOUT])
{
    const double* A1 = p;        // A1(i,j) = A1[i + 3j]
    const double* b1 = p + 6;
    const double* A2 = p + 9;    // A2(i,j) = A2[i + 3j]
    const double* b2 = p + 18;
    const double* A3 = p + 21;   // A3(0,j) = A3[j]
    const double b3 = p[24];
    const double* t32 = TAB ? t32s : nullptr;
    // ---- forward ----
    double x1[3], e1[3], y1[3], h[3], e2[3], s2[3], y2[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        x1[i] = (A1[i] * xa + A1[i + 3] * xb) + b1[i];          // dot(A1, xi) + b1
        e1[i] = exp_s(-x1[i], t32);
        y1[i] = inv1p(e1[i]);                            // sigmoid = 1/(1+exp(-x)), unary.hpp:273-275
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        h[i] = ((A2[i] * y1[0] + A2[i + 3] * y1[1]) + A2[i + 6] * y1[2]) + b2[i];
        e2[i] = exp_s(-h[i], t32);
        s2[i] = inv1p(e2[i]);
        y2[i] = s2[i] + x1[i];                                  // sigmoid(...) + x1
    }
    const double y3 = ((A3[0] * y2[0] + A3[1] * y2[1]) + A3[2] * y2[2]) + b3;
    const double d = yy - y3;
    const double loss = d * d;                                  // pow<2>
    acc[NP] += loss;
    // ---- backward, seed 1 ----
    const double g_d = 2. * (d == 0 ? 0 : loss / d);            // pow.hpp:125-140
    const double g_y3 = -g_d;                                   // Sub::brmap
    double g_y2[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        acc[21 + j] += g_y3 * y2[j];                            // A3_adj = adj * y2^T
        g_y2[j] = A3[j] * g_y3;                                 // A3^T * adj
    }
    acc[24] += g_y3;                                            // b3
    double g_h[3], g_y1[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        g_h[i] = g_y2[i] * e2[i] * (s2[i] * s2[i]);             // Sigmoid::bmap seed*e/((e+1)(e+1)), unary.hpp:276-278
        acc[18 + i] += g_h[i];                                  // b2
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            acc[9 + i + 3 * j] += g_h[i] * y1[j];               // A2_adj = adj * y1^T
            g_y1[j] += A2[i + 3 * j] * g_h[i];                  // A2^T * adj
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        // x1 enters twice (README form has the sub-tree twice): through y1 and the jump
        const double g_x1 = g_y1[i] * e1[i] * (y1[i] * y1[i]) + g_y2[i];
        acc[i] += g_x1 * xa;                                    // A1_adj = adj * xi^T
        acc[i + 3] += g_x1 * xb;
        acc[6 + i] += g_x1;                                     // b1
    }
}

// Warp-level reduce-scatter of the 26 channels by recursive halving: at the step with partner lane ^ H a lane keeps one half
// of its channels (which half: its own bit H) and hands the other half to the partner, so the five steps cost 16 + 8 + 4 + 2 + 1
// = 31 shuffles of a double instead of the 26 x 5 = 130 of one butterfly per channel (ncu before: the shuffles of the CTA
// reduction were ~2 us of the 10.6 us small-batch launch).  Afterwards lane l holds the warp's total of channel l (0 for
// l >= NOUT).  Fixed tree: the bits do not depend on timing.
__device__ __forceinline__ double warp_reduce_scatter26(const double (&acc)[NOUT], int lane)
{
    constexpr unsigned FULL = 0xffffffffu;
    double a16[16], a8[8], a4[4], a2[2];
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const double hi = (k + 16 < NOUT) ? acc[k + 16] : 0.0;
        const double keep = b4 ? hi : acc[k], give = b4 ? acc[k] : hi;
        a16[k] = keep + __shfl_xor_sync(FULL, give, 16);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double keep = b3 ? a16[k + 8] : a16[k], give = b3 ? a16[k] : a16[k + 8];
        a8[k] = keep + __shfl_xor_sync(FULL, give, 8);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double keep = b2 ? a8[k + 4] : a8[k], give = b2 ? a8[k] : a8[k + 4];
        a4[k] = keep + __shfl_xor_sync(FULL, give, 4);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const double keep = b1 ? a4[k + 2] : a4[k], give = b1 ? a4[k] : a4[k + 2];
        a2[k] = keep + __shfl_xor_sync(FULL, give, 2);
    }
    const double keep = b0 ? a2[1] : a2[0], give = b0 ? a2[0] : a2[1];
    return keep + __shfl_xor_sync(FULL, give, 1);
}

// ROWS rows per thread and trip (independent dependency chains for the FP64 pipe), MINB resident CTAs
// Link = P2PLink: batch shard of ONE evaluation over the ranks of a node; the last CTA pushes its 26 sums into every
// rank's mailbox over NVLink, waits for the other ranks and writes the rank-ordered totals (pass + all-reduce in one launch).
template <int ROWS, int MINB, int THREADS = 256, class Link = NoLink>
__global__ void __launch_bounds__(THREADS, MINB)
mlp3_kernel(size_t batch, const double* __restrict__ X, const double* __restrict__ y,
            const double* __restrict__ parm, double* cta_partials, unsigned int* ticket,
            double* grad, int overwrite, double* out_loss, Link link, int red_mode)
{
    pdl_prologue();
    constexpr bool TAB = MINB >= 2;          // large-batch variant: table exp (the copy costs the small-batch one more than it saves)
    __shared__ double p[NP];
    __shared__ double t32[TAB ? 32 : 1];
    __shared__ double red[THREADS / 32][32];
    __shared__ bool is_last;
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t nth = (size_t)gridDim.x * blockDim.x;
    // the first row's inputs are requested BEFORE the parameter copy and its barrier: one DRAM round trip instead of two
    // on the critical path of the small-batch launch (one row per thread, latency-bound)
    double xa0 = 0.0, xb0 = 0.0, yy0 = 0.0;
    if (ROWS == 1 && tid < batch) { xa0 = __ldg(X + tid); xb0 = __ldg(X + batch + tid); yy0 = __ldg(y + tid); }
    if (threadIdx.x < NP) p[threadIdx.x] = parm[threadIdx.x];
    if (TAB && threadIdx.x >= 32 && threadIdx.x < 64) t32[threadIdx.x - 32] = fm::kExp2_32_d[threadIdx.x - 32];
    __syncthreads();

    double acc[NOUT];
#pragma unroll
    for (int k = 0; k < NOUT; ++k) acc[k] = 0.0;

    if (ROWS == 1) {
        // software pipeline: the next row's three inputs are loaded before the current row is evaluated
        // (ncu before: long-scoreboard 2.1 cycles per issue at 4 warps per scheduler)
        size_t row = tid;
        if (row < batch) {
            double xa = xa0, xb = xb0, yy = yy0;
            while (true) {
                const size_t nxt = row + nth;
                const bool more = nxt < batch;
                double xa2 = 0.0, xb2 = 0.0, yy2 = 0.0;
                if (more) { xa2 = __ldg(X + nxt); xb2 = __ldg(X + batch + nxt); yy2 = __ldg(y + nxt); }
                mlp3_row<TAB>(p, t32, xa, xb, yy, acc);
                if (!more) break;
                row = nxt; xa = xa2; xb = xb2; yy = yy2;
            }
        }
    } else {
        for (size_t row0 = tid; row0 < batch; row0 += ROWS * nth) {
            double xa[ROWS], xb[ROWS], yy[ROWS];
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const size_t row = row0 + r * nth;
                const bool in = row < batch;
                xa[r] = in ? __ldg(X + row) : 0.0;
                xb[r] = in ? __ldg(X + batch + row) : 0.0;
                yy[r] = in ? __ldg(y + row) : 0.0;
            }
#pragma unroll
            for (int r = 0; r < ROWS; ++r)
                if (row0 + r * nth < batch) mlp3_row<TAB>(p, t32, xa[r], xb[r], yy[r], acc);
        }
    }
    // ---- CTA reduce of the 26 channels ----
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double v = 0.0;
    if (red_mode == 0) {
        // FADB_MLP_REDUCE=0, the round-1 protocol kept for A/B runs: one butterfly per channel, every warp fences twice,
        // 26 threads of the last CTA walk the partial rows
#pragma unroll
        for (int k = 0; k < NOUT; ++k) {
            const double w = warp_sum(acc[k]);
            if (lane == 0) red[warp][k] = w;
        }
        __syncthreads();
        if (threadIdx.x < NOUT) {
            double t = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w][threadIdx.x];
            cta_partials[(size_t)blockIdx.x * NOUT + threadIdx.x] = t;
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        __syncthreads();
        if (!is_last) return;
        __threadfence();
        if (threadIdx.x < NOUT)
            for (unsigned b = 0; b < gridDim.x; ++b) v += __ldcg(&cta_partials[(size_t)b * NOUT + threadIdx.x]);
    } else {
        red[warp][lane] = warp_reduce_scatter26(acc, lane);       // lane = channel
        __syncthreads();
        if (warp == 0) {
            // warp 0 alone publishes the CTA's row: its lanes store, the warp converges, ONE thread fences (cumulative over
            // the stores ordered before it by the warp barrier) and takes the ticket; the other 15 warps never fence
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < THREADS / 32; ++w) t += red[w][lane];
            if (lane < NOUT) cta_partials[(size_t)blockIdx.x * NOUT + lane] = t;
            __syncwarp();
            if (lane == 0) {
                __threadfence();
                is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
                if (is_last) __threadfence();                     // acquire side: the partial rows are read with ld.cg below
            }
        }
        __syncthreads();
        if (!is_last) return;
        // last CTA: warp w sums the rows of CTAs w, w + nwarps, ... (lane = channel), all of a lane's loads independent and in
        // flight together (one L2 round trip for up to 8 rows per warp); then the warp sums in warp order
        double t = 0.0;
        if (lane < NOUT) {
            unsigned b = warp;
            for (; b + 3 * (THREADS / 32) < gridDim.x; b += 4 * (THREADS / 32)) {
                const double t0 = __ldcg(&cta_partials[(size_t)b * NOUT + lane]);
                const double t1 = __ldcg(&cta_partials[(size_t)(b + THREADS / 32) * NOUT + lane]);
                const double t2 = __ldcg(&cta_partials[(size_t)(b + 2 * (THREADS / 32)) * NOUT + lane]);
                const double t3 = __ldcg(&cta_partials[(size_t)(b + 3 * (THREADS / 32)) * NOUT + lane]);
                t += (t0 + t1) + (t2 + t3);
            }
            for (; b < gridDim.x; b += THREADS / 32) t += __ldcg(&cta_partials[(size_t)b * NOUT + lane]);
        }
        __syncthreads();                                           // red[] is reused
        red[warp][lane] = t;
        __syncthreads();
        if (threadIdx.x < NOUT) {
#pragma unroll
            for (int w = 0; w < THREADS / 32; ++w) v += red[w][threadIdx.x];
        }
    }
    if constexpr (std::is_same<Link, P2PLink>::value) {
        p2p_push(link, (int)threadIdx.x, NOUT, v, [] { __syncthreads(); });
        v = p2p_pull(link, (int)threadIdx.x, NOUT);
    }
    if (threadIdx.x < NOUT) {
        if (threadIdx.x == NP) out_loss[0] = v;
        else if (grad) grad[threadIdx.x] = overwrite ? v : grad[threadIdx.x] + v;
    }
    if (threadIdx.x == 0) *ticket = 0u;
}

static double* g_mlp_ws[kMaxDevices] = {nullptr};
static ShutdownHook g_mlp_ws_release([](int d) { cudaFree(g_mlp_ws[d]); g_mlp_ws[d] = nullptr; });

}  // namespace fadb

using namespace fadb;

static int mlp3_launch(size_t batch, const double* X, const double* y, const double* parm, double* grad, unsigned flags,
                       double* dev_loss, bool exchange);

extern "C" int fadb_mlp3_async(size_t batch, const double* X, const double* y, const double* parm,
                               double* grad, unsigned flags, double* dev_loss)
{
    return mlp3_launch(batch, X, y, parm, grad, flags, dev_loss, false);
}

// This rank's rows with the 26-double all-reduce fused in (csrc/p2p.cu must be connected): grad / dev_loss receive the
// sums over ALL ranks' rows, identical bits on every rank; one launch, no collective call.
extern "C" int fadb_mlp3_allreduce_async(size_t batch, const double* X, const double* y, const double* parm,
                                         double* grad, unsigned flags, double* dev_loss)
{
    return mlp3_launch(batch, X, y, parm, grad, flags, dev_loss, true);
}

static int mlp3_launch(size_t batch, const double* X, const double* y, const double* parm, double* grad, unsigned flags,
                       double* dev_loss, bool exchange)
{
    FADB_REQUIRE(batch == 0 || (X && y), "fadb_mlp3: null data");
    FADB_REQUIRE(parm && dev_loss, "fadb_mlp3: null parameter / loss pointer");
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    if (!g_mlp_ws[c->device]) FADB_CUDA(cudaMalloc(&g_mlp_ws[c->device], sizeof(double) * kMaxGrid * NOUT));
    const int threads = 256;
    // FADB_MLP_REDUCE=0: the round-1 reduction protocol (A/B runs); default: reduce-scatter + single-thread fences + wide last pass
    static const int red = [] { const char* e = getenv("FADB_MLP_REDUCE"); return (e && e[0] == '0') ? 0 : 1; }();
    if (exchange) {
        // same variant choice as below, by batch size
        P2PLink link;
        if ((rc = p2p_next_link(*c, &link))) return rc;
        const int ow = (flags & FADB_OVERWRITE_ADJ) ? 1 : 0;
        if (batch > (size_t)4 * c->sms * threads) {
            auto K = mlp3_kernel<1, 2, 256, P2PLink>;
            FADB_CUDA(launch_pdl(K, resident_grid(*c, K, threads, 0, batch, threads), threads, 0, c->stream, batch, X, y, parm,
                                 g_mlp_ws[c->device], c->tickets + 5, grad, ow, dev_loss, link, red));
        } else if (batch > (size_t)c->sms * 256 && batch <= (size_t)c->sms * 512) {
            FADB_CUDA(launch_pdl(mlp3_kernel<1, 1, 512, P2PLink>, (int)std::min<size_t>((batch + 511) / 512, (size_t)c->sms), 512, 0,
                                 c->stream, batch, X, y, parm, g_mlp_ws[c->device], c->tickets + 5, grad, ow, dev_loss, link, red));
        } else {
            auto K = mlp3_kernel<1, 1, 256, P2PLink>;
            FADB_CUDA(launch_pdl(K, resident_grid(*c, K, threads, 0, batch, threads), threads, 0, c->stream, batch, X, y, parm,
                                 g_mlp_ws[c->device], c->tickets + 5, grad, ow, dev_loss, link, red));
        }
        c->launches++;
        FADB_CUDA(cudaGetLastError());
        return FADB_OK;
    }
    static const int variant = [] { const char* e = getenv("FADB_MLP_VARIANT"); return e ? atoi(e) : -2; }();
#define FADB_MLP_LAUNCH(K)                                                                                     \
    FADB_CUDA(launch_pdl(K, resident_grid(*c, K, threads, 0, batch, threads), threads, 0, c->stream, batch, X, y, parm, \
                         g_mlp_ws[c->device], c->tickets + 5, grad, (flags & FADB_OVERWRITE_ADJ) ? 1 : 0, dev_loss, NoLink{}, red))
    // measured on B200, batch 2^22: <1,1> 166 us (153 registers, 8 warps/SM: FP64 pipe 37 % active, stalls are
    // dependent-issue waits), <2,1> 155, <3,1> 159, <2,2> 121, <1,2> 115 (128 registers, 16 warps/SM): default;
    // <1,3> 152 and <1,4> 199 (80 / 64 registers: 430-530 bytes of spills per row)
    if (variant == 1) FADB_MLP_LAUNCH((mlp3_kernel<2, 1>));
    else if (variant == 3) FADB_MLP_LAUNCH((mlp3_kernel<2, 2>));
    else if (variant == 4) FADB_MLP_LAUNCH((mlp3_kernel<3, 1>));
    else if (variant == 0) FADB_MLP_LAUNCH((mlp3_kernel<1, 1>));
    else if (variant == 2) FADB_MLP_LAUNCH((mlp3_kernel<1, 2>));
    // default: occupancy wins once every CTA slot has work; below that the grid-wide reduction (one partial
    // row + one ticket per CTA) dominates and fewer CTAs are faster (2^16 rows: 11.1 us vs 13.9 us)
    else if (batch > (size_t)4 * c->sms * threads) FADB_MLP_LAUNCH((mlp3_kernel<1, 2>));
    else if (variant == 7 || (variant == -2 && batch > (size_t)c->sms * 256 && batch <= (size_t)c->sms * 512)) {
        // one row per thread in ONE trip with no more CTAs (= partial rows and tickets) than SMs: 512-thread CTAs
        FADB_CUDA(launch_pdl(mlp3_kernel<1, 1, 512>, (int)std::min<size_t>((batch + 511) / 512, (size_t)c->sms), 512, 0, c->stream, batch, X, y,
                             parm, g_mlp_ws[c->device], c->tickets + 5, grad, (flags & FADB_OVERWRITE_ADJ) ? 1 : 0, dev_loss, NoLink{}, red));
    }
    else FADB_MLP_LAUNCH((mlp3_kernel<1, 1>));
#undef FADB_MLP_LAUNCH
    c->launches++;
    FADB_CUDA(cudaGetLastError());
    return FADB_OK;
}

extern "C" int fadb_mlp3(size_t batch, const double* X, const double* y, const double* parm,
                         double* grad, unsigned flags, double* host_loss)
{
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    double* dval = c->scalars + 3;
    rc = fadb_mlp3_async(batch, X, y, parm, grad, flags, dval);
    if (rc) return rc;
    if (host_loss) {
        FADB_CUDA(cudaMemcpyAsync(c->host_scalars + 3, dval, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        FADB_CUDA(cudaStreamSynchronize(c->stream));
        *host_loss = c->host_scalars[3];
    }
    return FADB_OK;
}
