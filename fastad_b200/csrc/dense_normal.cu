// dense_normal.cu — K7: normal_adj_log_pdf(x, mu, Sigma) with a dense covariance matrix
// (SURVEY.md 8f row 3): NormalAdjLogPDFNode<vec, scl|vec, mat>::feval/beval,
// reverse/stat/normal.hpp:581-773 — Eigen::LLT<Lower> of Sigma, log_det = log det L,
// inv = llt.solve(I), z = inv (x - mu), value = -0.5 (x - mu)^T z - log_det, and the adjoints
// Sigma_adj += (-0.5 seed)(inv - z z^T), mu_adj += seed z (or its sum), x_adj += -seed z.
//
// All FP64, hand-written, deterministic (no atomics):
//   1. two-level right-looking Cholesky (128-wide outer panels, 32-wide inner steps): one warp
//      factors the diagonal tile in registers (and inverts it), a panel kernel applies L_kk^{-T},
//      a tiled SYRK updates the rest of the outer panel (rank 32) and the trailing matrix (rank 128);
//   2. Y = L^{-1} by recursive halving ([A 0; B C]^-1 = [A^-1 0; -C^-1 B A^-1, C^-1]), one batched launch
//      pair per tree level, on a tiled FP64 GEMM that skips the zero tiles of triangular operands;
//      X = Y^T Y on the lower tiles only
//      (a first version with one CTA per block column walking 32x32 products was 20x slower:
//      its critical path is nb^2/2 dependent tile products);
//   3. GEMV z = inv d, the value reduction and the adjoint scatter.
// The FP64 tensor path (DMMA m8n8k4) is not used: on B200 its peak equals the vector FP64 peak
// and the dependent 32-wide substitution chains, not the multiply throughput, bound this shape.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace fadb {

constexpr int DB = 32;

struct DenseWs {
    double *L = nullptr, *Y = nullptr, *X = nullptr, *Dinv = nullptr, *d = nullptr, *z = nullptr, *scal = nullptr;
    size_t cap = 0;
};
static DenseWs g_dws[kMaxDevices];

__device__ __forceinline__ double ld_pad(const double* A, int n, int r, int c)
{   // out-of-range entries behave like an identity border, so partial tiles need no special code
    return (r < n && c < n) ? A[r + (size_t)c * n] : (r == c ? 1.0 : 0.0);
}

__global__ void dn_copy_lower(int n, const double* __restrict__ S, double* L)
{
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n) return;
    const int i = (int)(idx % n), j = (int)(idx / n);
    L[idx] = i >= j ? S[idx] : 0.0;          // LLT<Lower> reads the lower triangle only
}

// diagonal tile: unblocked Cholesky + inverse of the 32x32 lower factor.  ONE warp, lane r keeps
// row r of the tile in registers; pivots and multipliers travel by shuffle, so there is no barrier
// and no shared memory on the 32-step dependent chain (a 32x32-thread version with block barriers
// took 31 us per tile, a shared-memory single-warp one 57 us; this one takes a few us).
// 1/sqrt(pivot) comes from the MUFU seed + Newton steps; the column is divided with one residual
// correction (within 1 ulp of the IEEE quotient).
__global__ void __launch_bounds__(32)
dn_potf2(int n, double* L, int k0, double* Dinv_k, double* flag)
{
    const int r = threadIdx.x;
    double row[DB], rinv[DB];
#pragma unroll
    for (int c = 0; c < DB; ++c) row[c] = (r >= c) ? ld_pad(L, n, k0 + r, k0 + c) : 0.0;
    bool bad = false;
#pragma unroll
    for (int j = 0; j < DB; ++j) {
        const double p = __shfl_sync(0xffffffffu, row[j], j);
        bad |= !(p > 0);                                    // llt.info() != Success, normal.hpp:660
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(p));
        const double h = 0.5 * p;
        y = y * fma(-h * y, y, 1.5);
        y = y * fma(-h * y, y, 1.5);
        double sq = p * y;
        sq = fma(fma(-sq, sq, p), 0.5 * y, sq);             // sqrt(p)
        y = fma(y, fma(-sq, y, 1.0), y);                    // 1 / sqrt(p)
        rinv[j] = y;
        double q = row[j] * y;
        q = fma(fma(-sq, q, row[j]), y, q);                 // row[j] / sqrt(p)
        row[j] = (r == j) ? sq : q;
#pragma unroll
        for (int c = j + 1; c < DB; ++c) {
            const double lcj = __shfl_sync(0xffffffffu, row[j], c);
            if (r >= c) row[c] = fma(-row[j], lcj, row[c]);
        }
    }
    if (bad && r == 0) flag[0] = 1.0;
#pragma unroll
    for (int c = 0; c < DB; ++c)
        if (c <= r && k0 + r < n && k0 + c < n) L[(k0 + r) + (size_t)(k0 + c) * n] = bad ? NAN : row[c];
    double inv[DB];                                         // lane r solves column r of T^{-1}
#pragma unroll
    for (int i = 0; i < DB; ++i) {
        double t0 = (i == r) ? 1.0 : 0.0, t1 = 0.0;
#pragma unroll
        for (int k = 0; k < i; ++k) {
            const double tik = __shfl_sync(0xffffffffu, row[k], i);
            if (k & 1) t1 = fma(-tik, inv[k], t1); else t0 = fma(-tik, inv[k], t0);
        }
        inv[i] = (i < r) ? 0.0 : (t0 + t1) * rinv[i];
    }
#pragma unroll
    for (int i = 0; i < DB; ++i) Dinv_k[i + r * DB] = bad ? NAN : inv[i];
}

// panel: P <- P * L_kk^{-T}, P = L[k0+32 : n, k0 : k0+32].  64 rows per CTA, 256 threads.
__global__ void __launch_bounds__(256)
dn_trsm_panel(int n, double* L, int k0, const double* __restrict__ Dinv_k)
{
    __shared__ double P[64][DB + 1], D[DB][DB + 1];
    const int t = threadIdx.x;
    const int r0 = k0 + DB + blockIdx.x * 64;
    for (int e = t; e < DB * DB; e += 256) D[e % DB][e / DB] = Dinv_k[e];
    for (int e = t; e < 64 * DB; e += 256) {
        const int rr = e % 64, cc = e / 64;
        P[rr][cc] = (r0 + rr < n && k0 + cc < n) ? L[(r0 + rr) + (size_t)(k0 + cc) * n] : 0.0;
    }
    __syncthreads();
    for (int e = t; e < 64 * DB; e += 256) {
        const int rr = e % 64, cc = e / 64;
        double acc = 0.0;
        for (int q = 0; q <= cc; ++q) acc += P[rr][q] * D[cc][q];      // (Dinv^T)[q][cc] = Dinv[cc][q]
        if (r0 + rr < n && k0 + cc < n) L[(r0 + rr) + (size_t)(k0 + cc) * n] = acc;
    }
}

// ---- general FP64 GEMM: C = alpha * op(A) * op(B) + beta * C, column-major, 64x64x16 tiles,
// 16x16 threads, 4x4 outputs per thread (rows along threadIdx.x so that stores coalesce).
// Triangular structure is exploited at tile granularity through `kmode`:
//   0 full K;  1 op(B) is lower triangular (K starts at the tile's first column);
//   2 op(A) is lower triangular (K ends at the tile's last row);
//   3 op(A) = M^T, op(B) = M with M lower triangular (K starts at max(tile row, tile col)).
// lower_only = 1 computes the tiles on or below the diagonal and masks the strict upper part.
template <bool TA, bool TB>
__device__ __forceinline__ void gemm_tile(int bi, int bj, int m, int n, int k, double alpha, const double* __restrict__ A,
                                          int lda, const double* __restrict__ B, int ldb, double beta, double* C, int ldc,
                                          int kmode, int lower_only)
{
    if (lower_only && bj > bi) return;
    __shared__ double As[16][64 + 4], Bs[16][64 + 4];
    const int tx = threadIdx.x, ty = threadIdx.y, t = ty * 16 + tx;
    const int i0 = bi * 64, j0 = bj * 64;
    if (i0 >= m || j0 >= n) return;
    int kb = 0, ke = k;
    if (kmode == 1) kb = j0;
    else if (kmode == 2) ke = min(k, i0 + 64);
    else if (kmode == 3) kb = max(i0, j0);
    kb &= ~15;
    double acc[4][4] = {};
    for (int k0 = kb; k0 < ke; k0 += 16) {
        for (int e = t; e < 64 * 16; e += 256) {
            int ii, kk;
            if (TA) { kk = e % 16; ii = e / 16; } else { ii = e % 64; kk = e / 64; }
            const int gi = i0 + ii, gk = k0 + kk;
            double v = 0.0;
            if (gi < m && gk < ke) v = TA ? A[gk + (size_t)gi * lda] : A[gi + (size_t)gk * lda];
            As[kk][ii] = v;
        }
        for (int e = t; e < 64 * 16; e += 256) {
            int jj, kk;
            if (TB) { jj = e % 64; kk = e / 64; } else { kk = e % 16; jj = e / 16; }
            const int gj = j0 + jj, gk = k0 + kk;
            double v = 0.0;
            if (gj < n && gk < ke) v = TB ? B[gj + (size_t)gk * ldb] : B[gk + (size_t)gj * ldb];
            Bs[kk][jj] = v;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            double a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { a[u] = As[q][tx * 4 + u]; b[u] = Bs[q][ty * 4 + u]; }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[u][v] = fma(a[u], b[v], acc[u][v]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + tx * 4 + u, j = j0 + ty * 4 + v;
            if (i < m && j < n && (!lower_only || i >= j)) {
                double* c = C + i + (size_t)j * ldc;
                *c = (beta == 0.0) ? alpha * acc[u][v] : fma(alpha, acc[u][v], beta * *c);
            }
        }
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
dn_gemm(int m, int n, int k, double alpha, const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
        double beta, double* C, int ldc, int kmode, int lower_only)
{
    gemm_tile<TA, TB>(blockIdx.x, blockIdx.y, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, kmode, lower_only);
}

// ---- batched triangular inversion.  Node z of the recursion tree owns W[off : off+n1+n2]^2 =
// [A 0; B C] with A^-1 (n1 x n1) and C^-1 (n2 x n2) already in place; all nodes of equal height are
// independent, so one launch handles a whole level:  phase 0: T_z = B A^-1,  phase 1: B <- -C^-1 T_z.
struct TrNode { int off, n1, n2, pad; long long toff; };

__global__ void __launch_bounds__(256)
dn_trtri_level(int n, double* W, double* T, const TrNode* __restrict__ nodes, int phase)
{
    const TrNode nd = nodes[blockIdx.z];
    double* Ainv = W + nd.off + (size_t)nd.off * n;
    double* Bm = W + (nd.off + nd.n1) + (size_t)nd.off * n;
    double* Cinv = W + (nd.off + nd.n1) + (size_t)(nd.off + nd.n1) * n;
    double* Tz = T + nd.toff;
    if (phase == 0)
        gemm_tile<false, false>(blockIdx.x, blockIdx.y, nd.n2, nd.n1, nd.n1, 1.0, Bm, n, Ainv, n, 0.0, Tz, nd.n2, 1, 0);
    else
        gemm_tile<false, false>(blockIdx.x, blockIdx.y, nd.n2, nd.n1, nd.n2, -1.0, Cinv, n, Tz, nd.n2, 0.0, Bm, n, 2, 0);
}

// copy the 32x32 inverse tiles of the diagonal blocks into W (start of the recursive inversion)
__global__ void dn_place_dinv(int n, int nb, const double* __restrict__ Dinv, double* W)
{
    const int b = blockIdx.x, r = threadIdx.x % DB, c = threadIdx.x / DB;
    for (int cc = c; cc < DB; cc += blockDim.x / DB) {
        const int gi = b * DB + r, gj = b * DB + cc;
        if (gi < n && gj < n) W[gi + (size_t)gj * n] = Dinv[(size_t)b * DB * DB + r + cc * DB];
    }
}
__global__ void dn_zero(size_t count, double* p)
{
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx < count) p[idx] = 0.0;
}

__global__ void dn_symmetrize(int n, double* X)
{
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n) return;
    const int i = (int)(idx % n), j = (int)(idx / n);
    if (i < j) X[idx] = X[j + (size_t)i * n];
}

// d = x - mu; z = inv d; one warp per row (inv is symmetric: row i == column i, contiguous)
__global__ void __launch_bounds__(256)
dn_gemv(int n, const double* __restrict__ X, const double* __restrict__ x, const double* __restrict__ mu, int mu_vec,
        double* d, double* z)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    const double m0 = mu_vec ? 0.0 : mu[0];
    double acc = 0.0;
    for (int k = lane; k < n; k += 32) acc += X[k + (size_t)warp * n] * (x[k] - (mu_vec ? mu[k] : m0));
    acc = warp_sum(acc);
    if (lane == 0) { z[warp] = acc; d[warp] = x[warp] - (mu_vec ? mu[warp] : m0); }
}

// value, scalar-mu adjoint and x / vector-mu adjoints.  1 CTA.
__global__ void __launch_bounds__(1024)
dn_finish(int n, const double* __restrict__ L, const double* __restrict__ d, const double* __restrict__ z,
          const double* flag, double seed, int overwrite, double* x_adj, double* mu_adj, int mu_vec, double* out_value)
{
    __shared__ double sm[32 * 3];
    double acc[3] = {0.0, 0.0, 0.0};             // d^T z, sum log L_ii, sum z
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        acc[0] += d[i] * z[i];
        acc[1] += log(L[i + (size_t)i * n]);
        acc[2] += z[i];
    }
    block_sum<3>(acc, sm);
    __shared__ bool ok;
    if (threadIdx.x == 0) {
        ok = flag[0] == 0.0;
        out_value[0] = ok ? -0.5 * acc[0] - acc[1] : -INFINITY;                    // normal.hpp:637-639
        if (ok && seed != 0.0 && mu_adj && !mu_vec) mu_adj[0] = overwrite ? seed * acc[2] : mu_adj[0] + seed * acc[2];
    }
    __syncthreads();
    if (!ok || seed == 0.0) return;                                                // normal.hpp:644
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (mu_adj && mu_vec) mu_adj[i] = overwrite ? seed * z[i] : mu_adj[i] + seed * z[i];
        if (x_adj) x_adj[i] = overwrite ? (-seed) * z[i] : x_adj[i] + (-seed) * z[i];
    }
}

__global__ void dn_sigma_adj(int n, const double* __restrict__ X, const double* __restrict__ z, const double* flag,
                             double seed, int overwrite, double* S_adj)
{
    if (flag[0] != 0.0) return;
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n) return;
    const int i = (int)(idx % n), j = (int)(idx / n);
    const double g = (-0.5 * seed) * (X[idx] - z[i] * z[j]);                       // normal.hpp:647-648
    S_adj[idx] = overwrite ? g : S_adj[idx] + g;
}

template <bool TA, bool TB>
static void gemm(Context& c, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb,
                 double beta, double* C, int ldc, int kmode, int lower_only)
{
    if (m <= 0 || n <= 0) return;
    dim3 grid((m + 63) / 64, (n + 63) / 64);
    dn_gemm<TA, TB><<<grid, dim3(16, 16), 0, c.stream>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, kmode, lower_only);
    c.launches++;
}

// recursion tree of the triangular inversion, grouped by height (leaves = single 32-wide tiles)
struct TrLevel { int first, count, max_n1, max_n2; };
struct TrPlan {
    size_t n = 0;
    std::vector<TrLevel> levels;
    TrNode* dev = nullptr;
};
static int tr_build(int off, int len, std::vector<std::vector<TrNode>>& by_height)
{
    if (len <= DB) return 0;
    const int n1 = std::max(DB, ((len / DB) / 2) * DB), n2 = len - n1;   // split on a tile boundary
    const int h = 1 + std::max(tr_build(off, n1, by_height), tr_build(off + n1, n2, by_height));
    if ((int)by_height.size() < h) by_height.resize(h);
    by_height[h - 1].push_back(TrNode{off, n1, n2, 0, 0});
    return h;
}
static TrPlan g_trplan[kMaxDevices];

static int ensure_trplan(Context& c, size_t n, TrPlan** out)
{
    TrPlan& p = g_trplan[c.device];
    if (p.n != n) {
        std::vector<std::vector<TrNode>> by_height;
        tr_build(0, (int)n, by_height);
        std::vector<TrNode> flat;
        p.levels.clear();
        for (auto& lv : by_height) {
            TrLevel L{(int)flat.size(), (int)lv.size(), 0, 0};
            long long toff = 0;                      // the T buffers of one level are packed into X
            for (auto& nd : lv) {
                nd.toff = toff;
                toff += (long long)nd.n1 * nd.n2;
                L.max_n1 = std::max(L.max_n1, nd.n1);
                L.max_n2 = std::max(L.max_n2, nd.n2);
                flat.push_back(nd);
            }
            p.levels.push_back(L);
        }
        FADB_CUDA(cudaStreamSynchronize(c.stream));
        if (p.dev) FADB_CUDA(cudaFree(p.dev));
        p.dev = nullptr;
        if (!flat.empty()) {
            FADB_CUDA(cudaMalloc(&p.dev, flat.size() * sizeof(TrNode)));
            FADB_CUDA(cudaMemcpy(p.dev, flat.data(), flat.size() * sizeof(TrNode), cudaMemcpyHostToDevice));
        }
        p.n = n;
    }
    *out = &p;
    return FADB_OK;
}

static int ensure_dws(Context& c, size_t n, DenseWs** out)
{
    DenseWs& w = g_dws[c.device];
    if (w.cap < n) {
        FADB_CUDA(cudaStreamSynchronize(c.stream));
        for (double* p : {w.L, w.Y, w.X, w.Dinv, w.d, w.z, w.scal}) if (p) FADB_CUDA(cudaFree(p));
        const size_t nb = (n + DB - 1) / DB;
        FADB_CUDA(cudaMalloc(&w.L, n * n * sizeof(double)));
        FADB_CUDA(cudaMalloc(&w.Y, n * n * sizeof(double)));
        FADB_CUDA(cudaMalloc(&w.X, n * n * sizeof(double)));
        FADB_CUDA(cudaMalloc(&w.Dinv, nb * DB * DB * sizeof(double)));
        FADB_CUDA(cudaMalloc(&w.d, n * sizeof(double)));
        FADB_CUDA(cudaMalloc(&w.z, n * sizeof(double)));
        FADB_CUDA(cudaMalloc(&w.scal, 8 * sizeof(double)));
        w.cap = n;
    }
    *out = &w;
    return FADB_OK;
}

}  // namespace fadb

using namespace fadb;

// factor = 1: (re)compute the Cholesky factor, log-det and inverse of Sigma; 0: reuse the ones of the
// previous call on this device (a constant Sigma is factored once, as normal.hpp:605-607 does)
extern "C" int fadb_normal_dense_async(size_t n, const double* x, const double* mu, int mu_vec, const double* Sigma,
                                       double* x_adj, double* mu_adj, double* Sigma_adj, double seed, unsigned flags,
                                       int factor, double* dev_value)
{
    FADB_REQUIRE(n >= 1 && n <= 46340, "fadb_normal_dense: n out of range");
    FADB_REQUIRE(x && mu && Sigma && dev_value, "fadb_normal_dense: null pointer");
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    DenseWs* w;
    rc = ensure_dws(*c, n, &w);
    if (rc) return rc;
    TrPlan* tp;
    rc = ensure_trplan(*c, n, &tp);
    if (rc) return rc;
    const int N = (int)n, nb = (N + DB - 1) / DB;
    const int ovw = (flags & FADB_OVERWRITE_ADJ) ? 1 : 0;
    const size_t nn = n * n;
    const int eb = (int)((nn + 255) / 256);
    cudaStream_t st = c->stream;
    if (factor) {
        FADB_CUDA(cudaMemsetAsync(w->scal, 0, 8 * sizeof(double), st));
        dn_copy_lower<<<eb, 256, 0, st>>>(N, Sigma, w->L);
        c->launches++;
        // two-level right-looking Cholesky: 128-wide outer panels, 32-wide inner steps.  The inner
        // rank-32 update touches the rest of the outer panel only; the trailing matrix gets one
        // rank-128 update per outer panel.
        constexpr int OB = 128;
        for (int K0 = 0; K0 < N; K0 += OB) {
            const int W = std::min(OB, N - K0), Kend = K0 + W;
            for (int k0 = K0; k0 < Kend; k0 += DB) {
                const int k = k0 / DB;
                dn_potf2<<<1, 32, 0, st>>>(N, w->L, k0, w->Dinv + (size_t)k * DB * DB, w->scal);
                c->launches++;
                const int m = N - k0 - DB;
                if (m <= 0) continue;
                dn_trsm_panel<<<(m + 63) / 64, 256, 0, st>>>(N, w->L, k0, w->Dinv + (size_t)k * DB * DB);
                c->launches++;
                const int pc = Kend - (k0 + DB);             // columns of the outer panel still to update
                if (pc > 0) {
                    const double* P = w->L + (k0 + DB) + (size_t)k0 * N;
                    double* Ct = w->L + (k0 + DB) + (size_t)(k0 + DB) * N;
                    gemm<false, true>(*c, m, pc, DB, -1.0, P, N, P, N, 1.0, Ct, N, 0, 1);
                }
            }
            const int m = N - Kend;
            if (m > 0) {
                const double* P = w->L + Kend + (size_t)K0 * N;
                double* Ct = w->L + Kend + (size_t)Kend * N;
                gemm<false, true>(*c, m, m, W, -1.0, P, N, P, N, 1.0, Ct, N, 0, 1);      // trailing -= P P^T (lower)
            }
        }
        // inverse: Y <- L^{-1} level by level up the halving tree, X <- Y^T Y (lower tiles, K from the diagonal)
        FADB_CUDA(cudaMemcpyAsync(w->Y, w->L, nn * sizeof(double), cudaMemcpyDeviceToDevice, st));
        dn_place_dinv<<<nb, 256, 0, st>>>(N, nb, w->Dinv, w->Y);
        c->launches++;
        for (const TrLevel& lv : tp->levels) {
            dim3 grid((lv.max_n2 + 63) / 64, (lv.max_n1 + 63) / 64, lv.count);
            dn_trtri_level<<<grid, dim3(16, 16), 0, st>>>(N, w->Y, w->X, tp->dev + lv.first, 0);
            dn_trtri_level<<<grid, dim3(16, 16), 0, st>>>(N, w->Y, w->X, tp->dev + lv.first, 1);
            c->launches += 2;
        }
        gemm<true, false>(*c, N, N, N, 1.0, w->Y, N, w->Y, N, 0.0, w->X, N, 3, 1);
        dn_symmetrize<<<eb, 256, 0, st>>>(N, w->X);
        c->launches++;
    }
    dn_gemv<<<(N * 32 + 255) / 256, 256, 0, st>>>(N, w->X, x, mu, mu_vec, w->d, w->z);
    dn_finish<<<1, 1024, 0, st>>>(N, w->L, w->d, w->z, w->scal, seed, ovw, x_adj, mu_adj, mu_vec, dev_value);
    c->launches += 2;
    if (Sigma_adj && seed != 0.0) {
        dn_sigma_adj<<<eb, 256, 0, st>>>(N, w->X, w->z, w->scal, seed, ovw, Sigma_adj);
        c->launches++;
    }
    FADB_CUDA(cudaGetLastError());
    return FADB_OK;
}

extern "C" int fadb_normal_dense(size_t n, const double* x, const double* mu, int mu_vec, const double* Sigma,
                                 double* x_adj, double* mu_adj, double* Sigma_adj, double seed, unsigned flags,
                                 double* host_value)
{
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    double* dval = c->scalars + 32;
    rc = fadb_normal_dense_async(n, x, mu, mu_vec, Sigma, x_adj, mu_adj, Sigma_adj, seed, flags, 1, dval);
    if (rc) return rc;
    if (host_value) {
        FADB_CUDA(cudaMemcpyAsync(c->host_scalars + 5, dval, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        FADB_CUDA(cudaStreamSynchronize(c->stream));
        *host_value = c->host_scalars[5];
    }
    return FADB_OK;
}
