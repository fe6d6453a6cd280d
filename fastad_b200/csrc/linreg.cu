// linreg.cu — K5: autodiff of normal_adj_log_pdf(y, dot(X, w), sigma) in one pass over X.
//
// Replaces DotNode::feval/beval (reverse/core/dot.hpp:89-104: Eigen GEMV forward, X^T*adj
// backward — X streamed twice) feeding NormalAdjLogPDFNode<vec,vec,scl>::feval/beval
// (reverse/stat/normal.hpp:338-372) with x = y constant, mean = dot(X, w), sigma a scalar
// variable, plus the leaf `adj += seed` of w and sigma (var_view.hpp:91-94).
//
// X is rows x cols column-major in HBM.  A CTA of 128 threads owns a tile of 64 rows; warp q
// owns columns [16q, 16q+16), lane l owns rows {2l, 2l+1} (one 128-bit load per column).  The
// 64x64 tile is held in REGISTERS between the forward product (t = X w) and the transposed
// product (X^T r), so X crosses HBM exactly once: 520 algorithmic bytes per row for cols=64.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "tma_util.cuh"
#include "p2p_device.cuh"

namespace fadb {

constexpr int LR_THREADS = 128;
constexpr int LR_ROWS = 64;      // rows per tile
constexpr int LR_CPW = 16;       // columns per warp per block of 64 columns

struct LinregArgs {
    size_t rows, cols;
    const double* X; const double* y; const double* w;
};

// partial layout per CTA: [0] = sum r^2, [1 .. cols] = (X^T r)_j
__global__ void __launch_bounds__(LR_THREADS)
linreg_kernel(LinregArgs a, double* cta_partials, unsigned int* ticket, double* out /* cols+1 */, P2PLink link)
{
    extern __shared__ double sm[];
    double* w_s = sm;                                  // [cols_pad]
    double* pt_s = w_s + ((a.cols + 63) / 64) * 64;    // [4][LR_ROWS]
    double* g_s = pt_s + 4 * LR_ROWS;                  // [cols_pad] cross-lane reduce target
    __shared__ bool is_last;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t cols_pad = ((a.cols + 63) / 64) * 64;
    const int ncb = (int)(cols_pad / 64);              // column blocks of 64
    for (size_t j = threadIdx.x; j < cols_pad; j += LR_THREADS) w_s[j] = j < a.cols ? a.w[j] : 0.0;
    __syncthreads();

    const bool vec_ok = ((reinterpret_cast<uintptr_t>(a.X) & 15u) == 0) && (a.rows % 2 == 0);
    const size_t ntiles = (a.rows + LR_ROWS - 1) / LR_ROWS;
    double rr = 0.0;                                   // sum r^2 (warp 0 only)
    // per-thread accumulators of (X^T r) for this warp's 16 columns of column block 0;
    // further column blocks (cols > 64) accumulate straight into shared memory atomically-free
    double g[LR_CPW];
#pragma unroll
    for (int c = 0; c < LR_CPW; ++c) g[c] = 0.0;
    for (size_t j = threadIdx.x; j < cols_pad; j += LR_THREADS) g_s[j] = 0.0;
    __syncthreads();

    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t r0 = tile * LR_ROWS + 2 * lane;   // this lane's two rows: r0, r0+1
        const bool in0 = r0 < a.rows, in1 = r0 + 1 < a.rows;
        double xv[LR_CPW][2];
        double t0 = 0.0, t1 = 0.0;
        // ---- forward: partial t over column block(s); block 0 stays in registers ----
        for (int cb = 0; cb < ncb; ++cb) {
            double xb[LR_CPW][2];
#pragma unroll
            for (int c = 0; c < LR_CPW; ++c) {
                const size_t j = (size_t)cb * 64 + warp * LR_CPW + c;
                double v0 = 0.0, v1 = 0.0;
                if (j < a.cols) {
                    const double* p = a.X + j * a.rows + r0;
                    if (vec_ok && in1) { const double2 v = __ldg(reinterpret_cast<const double2*>(p)); v0 = v.x; v1 = v.y; }
                    else { if (in0) v0 = __ldg(p); if (in1) v1 = __ldg(p + 1); }
                }
                xb[c][0] = v0; xb[c][1] = v1;
            }
            double s0a = 0.0, s0b = 0.0, s1a = 0.0, s1b = 0.0;
#pragma unroll
            for (int c = 0; c < LR_CPW; c += 2) {
                const double wa = w_s[cb * 64 + warp * LR_CPW + c], wb = w_s[cb * 64 + warp * LR_CPW + c + 1];
                s0a += xb[c][0] * wa; s0b += xb[c + 1][0] * wb;
                s1a += xb[c][1] * wa; s1b += xb[c + 1][1] * wb;
            }
            t0 += s0a + s0b; t1 += s1a + s1b;
            if (cb == 0) {
#pragma unroll
                for (int c = 0; c < LR_CPW; ++c) { xv[c][0] = xb[c][0]; xv[c][1] = xb[c][1]; }
            }
        }
        pt_s[warp * LR_ROWS + 2 * lane] = t0;
        pt_s[warp * LR_ROWS + 2 * lane + 1] = t1;
        __syncthreads();
        // ---- residual r = y - X w for this lane's two rows (each warp recomputes it) ----
        double r_0 = 0.0, r_1 = 0.0;
        {
            const double m0 = (pt_s[2 * lane] + pt_s[LR_ROWS + 2 * lane]) + (pt_s[2 * LR_ROWS + 2 * lane] + pt_s[3 * LR_ROWS + 2 * lane]);
            const double m1 = (pt_s[2 * lane + 1] + pt_s[LR_ROWS + 2 * lane + 1]) + (pt_s[2 * LR_ROWS + 2 * lane + 1] + pt_s[3 * LR_ROWS + 2 * lane + 1]);
            if (in0) r_0 = __ldg(a.y + r0) - m0;
            if (in1) r_1 = __ldg(a.y + r0 + 1) - m1;
        }
        if (warp == 0) rr += r_0 * r_0 + r_1 * r_1;
        // ---- transposed product: column block 0 from registers ----
#pragma unroll
        for (int c = 0; c < LR_CPW; ++c) g[c] += xv[c][0] * r_0 + xv[c][1] * r_1;
        // further column blocks: re-read X (L2-resident: the tile was just streamed)
        for (int cb = 1; cb < ncb; ++cb) {
#pragma unroll 4
            for (int c = 0; c < LR_CPW; ++c) {
                const size_t j = (size_t)cb * 64 + warp * LR_CPW + c;
                double v = 0.0;
                if (j < a.cols) {
                    const double* p = a.X + j * a.rows + r0;
                    if (in0) v += __ldg(p) * r_0;
                    if (in1) v += __ldg(p + 1) * r_1;
                }
                v = warp_sum(v);
                if (lane == 0) g_s[j] += v;     // one writer per column per CTA (warp owns j)
            }
        }
        __syncthreads();   // pt_s is reused by the next tile
    }
    // ---- CTA reduce: lanes -> column totals ----
#pragma unroll
    for (int c = 0; c < LR_CPW; ++c) {
        const double v = warp_sum(g[c]);
        if (lane == 0) g_s[warp * LR_CPW + c] += v;
    }
    rr = warp_sum(rr);
    __syncthreads();
    double* mine = cta_partials + (size_t)blockIdx.x * (a.cols + 1);
    if (threadIdx.x == 0) mine[0] = rr;
    for (size_t j = threadIdx.x; j < a.cols; j += LR_THREADS) mine[1 + j] = g_s[j];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // ---- last CTA: fixed-order sum over CTAs, one thread per output slot ----
    double mine_acc = 0.0;
    for (size_t j = threadIdx.x; j < a.cols + 1; j += LR_THREADS) {
        double acc = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) acc += __ldcg(&cta_partials[(size_t)b * (a.cols + 1) + j]);
        out[j] = acc;
        mine_acc = acc;                 // cols + 1 <= LR_THREADS whenever the exchange is fused: one slot per thread
    }
    if (threadIdx.x == 0) *ticket = 0u;
    // fused exchange: the reduced partials go straight to every rank's mailbox over NVLink (p2p_device.cuh)
    if (link.world) p2p_push(link, (int)threadIdx.x, (int)a.cols + 1, mine_acc, [] { __syncthreads(); });
}

// ---------------------------------------------------------------------------------------
// TMA variant (cols <= 64, even rows, 16-byte aligned X and y): a producer thread streams 64-row
// tiles of X (ONE cp.async.bulk.tensor.2d per 32 KB tile through a tensor map; per-column 512 B
// bulk copies were measured issue-bound) and of y into a 3-stage shared-memory ring guarded by
// full/empty mbarriers; four consumer warps do the two products.  Memory
// latency is hidden by the ring instead of by occupancy (the register kernel above sits at 214
// registers and 12 % occupancy), and each stage is released as soon as its tile is in registers.
// ---------------------------------------------------------------------------------------
constexpr int LT_STAGES = 3;
constexpr int LT_CONSUMERS = 128;
constexpr int LT_THREADS = LT_CONSUMERS + 32;
constexpr int LT_STAGE_BYTES = 64 * LR_ROWS * 8 + LR_ROWS * 8;   // X tile [64 cols][64 rows] + y tile

__device__ __forceinline__ void consumer_bar()
{
    asm volatile("bar.sync 1, %0;" ::"n"(LT_CONSUMERS) : "memory");
}

__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst_smem)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(LT_THREADS)
linreg_tma_kernel(LinregArgs a, const __grid_constant__ CUtensorMap xmap, int box_cols, double* cta_partials,
                  unsigned int* ticket, double* out /* cols+1 */, P2PLink link)
{
    pdl_prologue();
    extern __shared__ __align__(128) unsigned char lt_smem[];
    unsigned char* stages = lt_smem;                                       // [LT_STAGES][LT_STAGE_BYTES]
    double* w_s = reinterpret_cast<double*>(lt_smem + LT_STAGES * LT_STAGE_BYTES);   // [64]
    double* pt_s = w_s + 64;                                               // [2][4][LR_ROWS]
    double* g_s = pt_s + 2 * 4 * LR_ROWS;                                  // [64]
    uint64_t* full = reinterpret_cast<uint64_t*>(g_s + 64);                // [LT_STAGES]
    uint64_t* empty = full + LT_STAGES;                                    // [LT_STAGES]
    __shared__ bool is_last;
    __shared__ double rr_s;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cols = (int)a.cols;
    if (threadIdx.x < 64) w_s[threadIdx.x] = threadIdx.x < cols ? a.w[threadIdx.x] : 0.0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < LT_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const size_t ntiles = (a.rows + LR_ROWS - 1) / LR_ROWS;
    if (warp == 4) {
        // ------------------------------------------------ producer warp
        int stage = 0;
        unsigned phase = 0;
        for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            mbar_wait(&empty[stage], phase ^ 1);
            const size_t r0 = tile * LR_ROWS;
            const unsigned nrow = (unsigned)((a.rows - r0 < (size_t)LR_ROWS) ? a.rows - r0 : LR_ROWS);
            unsigned char* base = stages + (size_t)stage * LT_STAGE_BYTES;
            if (lane == 0) {
                // one TMA tensor copy per tile (box = 64 rows x box_cols columns, rows beyond the
                // matrix are zero-filled by the TMA unit) + one bulk copy for the y segment
                mbar_arrive_expect_tx(&full[stage], (unsigned)box_cols * LR_ROWS * 8u + nrow * 8u);
                tma_load_2d(base, &xmap, (int)r0, 0, &full[stage]);
                bulk_g2s(base + 64 * LR_ROWS * 8, a.y + r0, nrow * 8u, &full[stage]);
            }
            if (++stage == LT_STAGES) { stage = 0; phase ^= 1; }
        }
        return;
    }
    // ---------------------------------------------------- consumer warps
    double g[LR_CPW];
#pragma unroll
    for (int c = 0; c < LR_CPW; ++c) g[c] = 0.0;
    double rr = 0.0;
    int stage = 0;
    unsigned phase = 0, par = 0;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t r0 = tile * LR_ROWS + 2 * lane;
        const bool in0 = r0 < a.rows, in1 = r0 + 1 < a.rows;
        mbar_wait(&full[stage], phase);
        const unsigned char* base = stages + (size_t)stage * LT_STAGE_BYTES;
        double2 xv[LR_CPW];
#pragma unroll
        for (int c = 0; c < LR_CPW; ++c) {
            const int j = warp * LR_CPW + c;
            double2 v = make_double2(0.0, 0.0);
            if (j < cols) v = *reinterpret_cast<const double2*>(base + (size_t)j * LR_ROWS * 8 + lane * 16);
            if (!in0) v.x = 0.0;
            if (!in1) v.y = 0.0;
            xv[c] = v;
        }
        const double2 yv = *reinterpret_cast<const double2*>(base + 64 * LR_ROWS * 8 + lane * 16);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);     // the tile is in registers: free the slot
        if (++stage == LT_STAGES) { stage = 0; phase ^= 1; }

        double s0a = 0.0, s0b = 0.0, s1a = 0.0, s1b = 0.0;
#pragma unroll
        for (int c = 0; c < LR_CPW; c += 2) {
            const double wa = w_s[warp * LR_CPW + c], wb = w_s[warp * LR_CPW + c + 1];
            s0a += xv[c].x * wa; s0b += xv[c + 1].x * wb;
            s1a += xv[c].y * wa; s1b += xv[c + 1].y * wb;
        }
        double* pt = pt_s + par * 4 * LR_ROWS;
        *reinterpret_cast<double2*>(pt + warp * LR_ROWS + 2 * lane) = make_double2(s0a + s0b, s1a + s1b);
        consumer_bar();
        const double2 p0 = *reinterpret_cast<const double2*>(pt + 2 * lane);
        const double2 p1 = *reinterpret_cast<const double2*>(pt + LR_ROWS + 2 * lane);
        const double2 p2 = *reinterpret_cast<const double2*>(pt + 2 * LR_ROWS + 2 * lane);
        const double2 p3 = *reinterpret_cast<const double2*>(pt + 3 * LR_ROWS + 2 * lane);
        par ^= 1;
        const double r_0 = in0 ? yv.x - ((p0.x + p1.x) + (p2.x + p3.x)) : 0.0;
        const double r_1 = in1 ? yv.y - ((p0.y + p1.y) + (p2.y + p3.y)) : 0.0;
        if (warp == 0) rr += r_0 * r_0 + r_1 * r_1;
#pragma unroll
        for (int c = 0; c < LR_CPW; ++c) g[c] += xv[c].x * r_0 + xv[c].y * r_1;
    }
    // ---- CTA reduce: lanes -> column totals (consumer threads only from here on)
#pragma unroll
    for (int c = 0; c < LR_CPW; ++c) {
        const double v = warp_sum(g[c]);
        if (lane == 0) g_s[warp * LR_CPW + c] = v;
    }
    rr = warp_sum(rr);
    if (threadIdx.x == 0) rr_s = rr;
    consumer_bar();
    double* mine = cta_partials + (size_t)blockIdx.x * (a.cols + 1);
    if (threadIdx.x == 0) mine[0] = rr_s;
    if ((int)threadIdx.x < cols) mine[1 + threadIdx.x] = g_s[threadIdx.x];
    consumer_bar();
    // ONE thread publishes the row: its fence is cumulative over the stores the barrier ordered before it
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        if (is_last) __threadfence();             // acquire side; the rows are read with ld.cg below
    }
    consumer_bar();
    if (!is_last) return;
    // last CTA: warp q sums rows q, q + 4, ... (lane = slot, slot + 32, slot + 64), eight rows' loads in flight per lane,
    // then the four warp sums in warp order (the 65 single-thread walks over 296 rows this replaces cost ~6 us)
    const int nslot = cols + 1;
    double acc3[3] = {0.0, 0.0, 0.0};
    {
        const size_t stride = a.cols + 1;
        unsigned b = warp;
        for (; b + 28 < gridDim.x; b += 32) {
            double v[8][3];
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int slot = lane + 32 * k;
                    v[u][k] = slot < nslot ? __ldcg(&cta_partials[(size_t)(b + 4 * u) * stride + slot]) : 0.0;
                }
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int k = 0; k < 3; ++k) acc3[k] += v[u][k];
        }
        for (; b < gridDim.x; b += 4)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int slot = lane + 32 * k;
                if (slot < nslot) acc3[k] += __ldcg(&cta_partials[(size_t)b * stride + slot]);
            }
    }
    double* red = pt_s;                           // [4][96] of the 512 doubles (the tile loop is over)
#pragma unroll
    for (int k = 0; k < 3; ++k) red[warp * 96 + lane + 32 * k] = acc3[k];
    consumer_bar();
    double mine_acc = 0.0;
    if ((int)threadIdx.x < nslot) {
        mine_acc = ((red[threadIdx.x] + red[96 + threadIdx.x]) + red[192 + threadIdx.x]) + red[288 + threadIdx.x];
        out[threadIdx.x] = mine_acc;
    }
    if (threadIdx.x == 0) *ticket = 0u;
    // fused exchange: the reduced partials go straight to every rank's mailbox over NVLink (p2p_device.cuh)
    if (link.world) p2p_push(link, (int)threadIdx.x, cols + 1, mine_acc, [] { consumer_bar(); });
}

// scalar epilogue: value + w_adj + sigma_adj from {sum r^2, X^T r} (normal.hpp:350-371, dot.hpp:96-104)
__global__ void linreg_finish_kernel(double n_total, size_t cols, const double* partial,
                                     const double* sigma, double* w_adj, double* sigma_adj,
                                     double seed, int overwrite, double* out_value, P2PLink link)
{
    __shared__ double s_part[P2P_MAXN];
    if (link.world) {       // fused exchange: wait for every rank's push, reduce in rank order (cols + 1 <= 128)
        const double v = p2p_pull(link, (int)threadIdx.x, (int)cols + 1);
        if (threadIdx.x < cols + 1) s_part[threadIdx.x] = v;
        __syncthreads();
        partial = s_part;
    }
    const double s = sigma[0];
    if (!(s > 0)) {
        if (threadIdx.x == 0) out_value[0] = -INFINITY;       // normal.hpp:344
        return;
    }
    const double inv_s = 1. / s, inv_s_sq = inv_s * inv_s;
    const double z_sq = partial[0] / (s * s);
    if (threadIdx.x == 0) {
        out_value[0] = -0.5 * z_sq - n_total * log(s);
        if (seed != 0.0 && sigma_adj) {
            const double g = seed * (z_sq - n_total) * inv_s;
            sigma_adj[0] = overwrite ? g : sigma_adj[0] + g;
        }
    }
    if (seed == 0.0 || !w_adj) return;
    const double c = seed * inv_s_sq;
    for (size_t j = threadIdx.x; j < cols; j += blockDim.x) {
        const double g = c * partial[1 + j];
        w_adj[j] = overwrite ? g : w_adj[j] + g;
    }
}

// least squares on the same partials: value = sum r^2, w_adj (+)= seed * (-2) * X^T r  with r = y - X w
// (pow.hpp:125-140 through Sub::brmap, binary.hpp:296-303, into dot.hpp:96-104)
__global__ void lsq_finish_kernel(size_t cols, const double* partial, double* w_adj, double seed, int overwrite,
                                  double* out_value)
{
    if (threadIdx.x == 0) out_value[0] = partial[0];
    if (!w_adj) return;
    const double c = -2. * seed;
    for (size_t j = threadIdx.x; j < cols; j += blockDim.x) {
        const double g = c * partial[1 + j];
        w_adj[j] = overwrite ? g : w_adj[j] + g;
    }
}

// rank-2 tensor map over the column-major X: dim0 = rows (contiguous), dim1 = cols
static int encode_xmap(CUtensorMap* map, const double* X, size_t rows, size_t cols)
{
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        FADB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess) {
            set_error("fastad_b200: cuTensorMapEncodeTiled is not available from this driver");
            return FADB_ERR_CUDA;
        }
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
    const cuuint64_t gstride[1] = {(cuuint64_t)rows * sizeof(double)};
    const cuuint32_t box[2] = {(cuuint32_t)LR_ROWS, (cuuint32_t)cols};
    const cuuint32_t estride[2] = {1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(X), gdim, gstride, box, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("fastad_b200: cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
        return FADB_ERR_CUDA;
    }
    return FADB_OK;
}

// for expr.cu's one-pass dot -> map -> sum kernel (glm_kernel.cuh): same tile geometry
int linreg_encode_xmap(void* map128, const double* X, size_t rows, size_t cols)
{
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
    return encode_xmap(static_cast<CUtensorMap*>(map128), X, rows, cols);
}

static double* g_ws[kMaxDevices] = {nullptr};
static size_t g_ws_cap[kMaxDevices] = {0};
static ShutdownHook g_ws_release([](int d) { cudaFree(g_ws[d]); g_ws[d] = nullptr; g_ws_cap[d] = 0; });

static int ensure_ws(Context& c, size_t doubles, double** out)
{
    if (g_ws_cap[c.device] < doubles) {
        if (g_ws[c.device]) { FADB_CUDA(cudaStreamSynchronize(c.stream)); FADB_CUDA(cudaFree(g_ws[c.device])); }
        FADB_CUDA(cudaMalloc(&g_ws[c.device], doubles * sizeof(double)));
        g_ws_cap[c.device] = doubles;
    }
    *out = g_ws[c.device];
    return FADB_OK;
}

}  // namespace fadb

using namespace fadb;

// dev_partial: cols+1 device doubles = {sum r^2, (X^T r)_0..}; all-reduce them across row
// shards, then call fadb_linreg_finish with the global row count.
static int linreg_partial_impl(size_t rows, size_t cols, const double* X, const double* y, const double* w, double* dev_partial,
                               bool push)
{
    FADB_REQUIRE(cols >= 1 && cols <= 4096, "fadb_linreg: cols must be in 1..4096");
    FADB_REQUIRE(X && y && w && dev_partial, "fadb_linreg: null pointer");
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    P2PLink link;
    link.world = 0;
    if (push) {
        FADB_REQUIRE(cols + 1 <= (size_t)P2P_MAXN, "fadb_linreg_partial_push: the fused exchange carries at most 128 doubles (cols <= 127)");
        if ((rc = p2p_next_link(*c, &link))) return rc;
    }
    const size_t ntiles = (rows + LR_ROWS - 1) / LR_ROWS;
    int grid = (int)std::min<size_t>(std::max<size_t>(ntiles, 1), (size_t)c->sms * 4);
    double* ws;
    rc = ensure_ws(*c, (size_t)grid * (cols + 1) + cols + 1, &ws);
    if (rc) return rc;
    static const bool no_tma = getenv("FADB_LINREG_NO_TMA") != nullptr;
    const bool tma_ok = !no_tma && cols <= 64 && rows % 2 == 0 && rows >= LR_ROWS &&
                        ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(y)) & 15u) == 0;
    if (tma_ok) {
        const size_t smem = (size_t)LT_STAGES * LT_STAGE_BYTES + (64 + 2 * 4 * LR_ROWS + 64) * sizeof(double) +
                            2 * LT_STAGES * sizeof(uint64_t);
        static bool attr_set[kMaxDevices] = {false};
        if (!attr_set[c->device]) {
            FADB_CUDA(cudaFuncSetAttribute(linreg_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set[c->device] = true;
        }
        int per_sm = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, linreg_tma_kernel, LT_THREADS, smem);
        if (per_sm < 1) per_sm = 1;
        const int g2 = (int)std::min<size_t>(ntiles, (size_t)c->sms * std::min(per_sm, 4));
        LinregArgs a{rows, cols, X, y, w};
        CUtensorMap xmap;
        rc = encode_xmap(&xmap, X, rows, cols);
        if (rc) return rc;
        FADB_CUDA(launch_pdl(linreg_tma_kernel, g2, LT_THREADS, smem, c->stream, a, xmap, (int)cols, ws, c->tickets + 4, dev_partial, link));
        c->launches++;
        FADB_CUDA(cudaGetLastError());
        return FADB_OK;
    }
    const size_t cols_pad = ((cols + 63) / 64) * 64;
    const size_t smem = (2 * cols_pad + 4 * LR_ROWS) * sizeof(double);
    if (smem > 48 * 1024)
        FADB_CUDA(cudaFuncSetAttribute(linreg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LinregArgs a{rows, cols, X, y, w};
    linreg_kernel<<<grid, LR_THREADS, smem, c->stream>>>(a, ws, c->tickets + 4, dev_partial, link);
    c->launches++;
    FADB_CUDA(cudaGetLastError());
    return FADB_OK;
}

extern "C" int fadb_linreg_partial(size_t rows, size_t cols, const double* X, const double* y,
                                   const double* w, double* dev_partial)
{
    return linreg_partial_impl(rows, cols, X, y, w, dev_partial, false);
}

// The same pass with the exchange FUSED into it: the CTA that forms {sum r^2, X^T r} stores them straight into
// every rank's mailbox over NVLink (csrc/p2p.cu; cols <= 127).  Pair with fadb_linreg_finish_pull: no
// collective call and no extra launch between the two kernels.
extern "C" int fadb_linreg_partial_push(size_t rows, size_t cols, const double* X, const double* y,
                                        const double* w, double* dev_partial)
{
    return linreg_partial_impl(rows, cols, X, y, w, dev_partial, true);
}

static int linreg_finish_impl(size_t rows_total, size_t cols, const double* dev_partial, const double* sigma, double* w_adj,
                              double* sigma_adj, double seed, unsigned flags, double* host_value, bool pull)
{
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    P2PLink link;
    link.world = 0;
    if (pull && (rc = p2p_current_link(*c, &link))) return rc;
    double* dval = c->scalars + 2;
    linreg_finish_kernel<<<1, 128, 0, c->stream>>>((double)rows_total, cols, dev_partial, sigma, w_adj,
                                                   sigma_adj, seed, (flags & FADB_OVERWRITE_ADJ) ? 1 : 0, dval, link);
    c->launches++;
    FADB_CUDA(cudaGetLastError());
    if (host_value) {
        FADB_CUDA(cudaMemcpyAsync(c->host_scalars + 2, dval, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        FADB_CUDA(cudaStreamSynchronize(c->stream));
        *host_value = c->host_scalars[2];
    }
    return FADB_OK;
}

extern "C" int fadb_linreg_finish(size_t rows_total, size_t cols, const double* dev_partial,
                                  const double* sigma, double* w_adj, double* sigma_adj, double seed,
                                  unsigned flags, double* host_value)
{
    return linreg_finish_impl(rows_total, cols, dev_partial, sigma, w_adj, sigma_adj, seed, flags, host_value, false);
}

// Consumer of fadb_linreg_partial_push: waits for every rank's push, sums the mailbox slots in rank order and
// forms the value and the adjoints, all in the one finish kernel.
extern "C" int fadb_linreg_finish_pull(size_t rows_total, size_t cols, const double* sigma, double* w_adj, double* sigma_adj,
                                       double seed, unsigned flags, double* host_value)
{
    FADB_REQUIRE(cols + 1 <= (size_t)P2P_MAXN, "fadb_linreg_finish_pull: cols <= 127");
    return linreg_finish_impl(rows_total, cols, nullptr, sigma, w_adj, sigma_adj, seed, flags, host_value, true);
}

// Least squares sum((y - X w)^2), value + gradient from ONE pass over X: the batched form of the
// reference's README regression example (README.md:606-662).  dev_value: device double.
extern "C" int fadb_lsq_finish(size_t cols, const double* dev_partial, double* w_adj, double seed, unsigned flags,
                               double* dev_value)
{
    FADB_REQUIRE(dev_partial && dev_value, "fadb_lsq_finish: null argument");
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    lsq_finish_kernel<<<1, 128, 0, c->stream>>>(cols, dev_partial, w_adj, seed, (flags & FADB_OVERWRITE_ADJ) ? 1 : 0, dev_value);
    c->launches++;
    FADB_CUDA(cudaGetLastError());
    return FADB_OK;
}

extern "C" int fadb_lsq(size_t rows, size_t cols, const double* X, const double* y, const double* w, double* w_adj,
                        double seed, unsigned flags, double* host_value)
{
    FADB_REQUIRE(cols >= 1 && cols <= 4096, "fadb_lsq: cols must be in 1..4096");
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    const size_t ntiles = (rows + LR_ROWS - 1) / LR_ROWS;
    const size_t grid = std::min<size_t>(std::max<size_t>(ntiles, 1), (size_t)c->sms * 4);
    double* ws;
    rc = ensure_ws(*c, grid * (cols + 1) + cols + 1, &ws);
    if (rc) return rc;
    double* partial = ws + grid * (cols + 1);
    rc = fadb_linreg_partial(rows, cols, X, y, w, partial);
    if (rc) return rc;
    double* dval = c->scalars + 40;
    rc = fadb_lsq_finish(cols, partial, w_adj, seed, flags, dval);
    if (rc) return rc;
    if (host_value) {
        FADB_CUDA(cudaMemcpyAsync(c->host_scalars + 40, dval, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        FADB_CUDA(cudaStreamSynchronize(c->stream));
        *host_value = c->host_scalars[40];
    }
    return FADB_OK;
}

extern "C" int fadb_linreg_nll(size_t rows, size_t cols, const double* X, const double* y,
                               const double* w, double* w_adj, const double* sigma,
                               double* sigma_adj, double seed, unsigned flags, double* host_value)
{
    FADB_REQUIRE(sigma != nullptr, "fadb_linreg_nll: null sigma");
    FADB_REQUIRE(cols >= 1 && cols <= 4096, "fadb_linreg: cols must be in 1..4096");
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    const size_t ntiles = (rows + LR_ROWS - 1) / LR_ROWS;
    const size_t grid = std::min<size_t>(std::max<size_t>(ntiles, 1), (size_t)c->sms * 4);
    double* ws;
    rc = ensure_ws(*c, grid * (cols + 1) + cols + 1, &ws);
    if (rc) return rc;
    double* partial = ws + grid * (cols + 1);
    rc = fadb_linreg_partial(rows, cols, X, y, w, partial);
    if (rc) return rc;
    return fadb_linreg_finish(rows, cols, partial, sigma, w_adj, sigma_adj, seed, flags, host_value);
}
