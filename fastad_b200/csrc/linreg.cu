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
#include <algorithm>

#include "common.cuh"

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
linreg_kernel(LinregArgs a, double* cta_partials, unsigned int* ticket, double* out /* cols+1 */)
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
    for (size_t j = threadIdx.x; j < a.cols + 1; j += LR_THREADS) {
        double acc = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) acc += __ldcg(&cta_partials[(size_t)b * (a.cols + 1) + j]);
        out[j] = acc;
    }
    if (threadIdx.x == 0) *ticket = 0u;
}

// scalar epilogue: value + w_adj + sigma_adj from {sum r^2, X^T r} (normal.hpp:350-371, dot.hpp:96-104)
__global__ void linreg_finish_kernel(double n_total, size_t cols, const double* partial,
                                     const double* sigma, double* w_adj, double* sigma_adj,
                                     double seed, int overwrite, double* out_value)
{
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

static double* g_ws[kMaxDevices] = {nullptr};
static size_t g_ws_cap[kMaxDevices] = {0};

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
extern "C" int fadb_linreg_partial(size_t rows, size_t cols, const double* X, const double* y,
                                   const double* w, double* dev_partial)
{
    FADB_REQUIRE(cols >= 1 && cols <= 4096, "fadb_linreg: cols must be in 1..4096");
    FADB_REQUIRE(X && y && w && dev_partial, "fadb_linreg: null pointer");
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    const size_t ntiles = (rows + LR_ROWS - 1) / LR_ROWS;
    int grid = (int)std::min<size_t>(std::max<size_t>(ntiles, 1), (size_t)c->sms * 4);
    double* ws;
    rc = ensure_ws(*c, (size_t)grid * (cols + 1) + cols + 1, &ws);
    if (rc) return rc;
    const size_t cols_pad = ((cols + 63) / 64) * 64;
    const size_t smem = (2 * cols_pad + 4 * LR_ROWS) * sizeof(double);
    if (smem > 48 * 1024)
        FADB_CUDA(cudaFuncSetAttribute(linreg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LinregArgs a{rows, cols, X, y, w};
    linreg_kernel<<<grid, LR_THREADS, smem, c->stream>>>(a, ws, c->tickets + 4, dev_partial);
    c->launches++;
    FADB_CUDA(cudaGetLastError());
    return FADB_OK;
}

extern "C" int fadb_linreg_finish(size_t rows_total, size_t cols, const double* dev_partial,
                                  const double* sigma, double* w_adj, double* sigma_adj, double seed,
                                  unsigned flags, double* host_value)
{
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    double* dval = c->scalars + 2;
    linreg_finish_kernel<<<1, 128, 0, c->stream>>>((double)rows_total, cols, dev_partial, sigma, w_adj,
                                                   sigma_adj, seed, (flags & FADB_OVERWRITE_ADJ) ? 1 : 0, dval);
    c->launches++;
    FADB_CUDA(cudaGetLastError());
    if (host_value) {
        FADB_CUDA(cudaMemcpyAsync(c->host_scalars + 2, dval, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        FADB_CUDA(cudaStreamSynchronize(c->stream));
        *host_value = c->host_scalars[2];
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
