// dense_normal.cu — K7: normal_adj_log_pdf(x, mu, Sigma) with a dense covariance matrix
// (SURVEY.md 8f row 3): NormalAdjLogPDFNode<vec, scl|vec, mat>::feval/beval,
// reverse/stat/normal.hpp:581-773 — Eigen::LLT<Lower> of Sigma, log_det = log det L,
// inv = llt.solve(I), z = inv (x - mu), value = -0.5 (x - mu)^T z - log_det, and the adjoints
// Sigma_adj += (-0.5 seed)(inv - z z^T), mu_adj += seed z (or its sum), x_adj += -seed z.
//
// All FP64, hand-written, deterministic (no atomics):
//   1. blocked right-looking Cholesky, block 32: one CTA factors the diagonal tile (and inverts
//      it), a panel kernel applies L_kk^{-T}, a tiled SYRK updates the trailing lower triangle;
//   2. Y = L^{-1} and X = L^{-T} Y by block-column-parallel substitution: one CTA owns a 32-wide
//      column of the right-hand side and walks the block rows with 32x32x32 register-tiled
//      products (4x4 outputs per thread);
//   3. GEMV z = inv d, the value reduction and the adjoint scatter.
// The FP64 tensor path (DMMA m8n8k4) is not used: on B200 its peak equals the vector FP64 peak
// and the dependent 32-wide substitution chains, not the multiply throughput, bound this shape.
#include <algorithm>

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

// diagonal tile: unblocked Cholesky + inverse of the 32x32 lower factor.  1 CTA, 32x32 threads.
__global__ void __launch_bounds__(1024)
dn_potf2(int n, double* L, int k0, double* Dinv_k, double* flag)
{
    __shared__ double T[DB][DB + 1], Inv[DB][DB + 1];
    const int r = threadIdx.y, c = threadIdx.x;
    T[r][c] = (r >= c) ? ld_pad(L, n, k0 + r, k0 + c) : 0.0;
    __syncthreads();
    for (int j = 0; j < DB; ++j) {
        if (r == j && c == j) {
            double p = T[j][j];
            if (!(p > 0)) { flag[0] = 1.0; p = NAN; }      // llt.info() != Success, normal.hpp:660
            T[j][j] = sqrt(p);
        }
        __syncthreads();
        if (c == j && r > j) T[r][j] /= T[j][j];
        __syncthreads();
        if (r > j && c > j && r >= c) T[r][c] -= T[r][j] * T[c][j];
        __syncthreads();
    }
    if (k0 + r < n && k0 + c < n && r >= c) L[(k0 + r) + (size_t)(k0 + c) * n] = T[r][c];
    if (r == 0) {                                  // thread c solves column c of T^{-1}
        for (int i = 0; i < DB; ++i) {
            double t = (i == c) ? 1.0 : 0.0;
            for (int k = c; k < i; ++k) t -= T[i][k] * Inv[k][c];
            Inv[i][c] = (i < c) ? 0.0 : t / T[i][i];
        }
    }
    __syncthreads();
    Dinv_k[r + c * DB] = Inv[r][c];
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

// trailing update: C[i][j] -= sum_q P[i][q] P[j][q] on 64x64 tiles with bi >= bj.  16x16 threads, 4x4 each.
__global__ void __launch_bounds__(256)
dn_syrk(int n, double* L, int k0)
{
    if (blockIdx.x > blockIdx.y) return;           // x = tile column, y = tile row: lower tiles only
    __shared__ double Pi[DB][64 + 4], Pj[DB][64 + 4];
    const int base = k0 + DB;
    const int i0 = base + blockIdx.y * 64, j0 = base + blockIdx.x * 64;
    const int t = threadIdx.y * 16 + threadIdx.x;
    for (int e = t; e < 64 * DB; e += 256) {
        const int rr = e % 64, q = e / 64;
        Pi[q][rr] = (i0 + rr < n) ? L[(i0 + rr) + (size_t)(k0 + q) * n] : 0.0;
        Pj[q][rr] = (j0 + rr < n) ? L[(j0 + rr) + (size_t)(k0 + q) * n] : 0.0;
    }
    __syncthreads();
    double acc[4][4] = {};
#pragma unroll 8
    for (int q = 0; q < DB; ++q) {
        double a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { a[u] = Pi[q][threadIdx.y * 4 + u]; b[u] = Pj[q][threadIdx.x * 4 + u]; }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[u][v] += a[u] * b[v];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int i = i0 + threadIdx.y * 4 + u, j = j0 + threadIdx.x * 4 + v;
            if (i < n && j < n && i >= j) L[i + (size_t)j * n] -= acc[u][v];
        }
}

// ---- 32x32x32 products on a CTA of 64 threads (8x8), 4x4 outputs per thread -----------------
// As is stored [q][row] (A^T) and Bs [q][col] so that both operand fetches are 4 consecutive doubles.
struct Tile { double v[4][4]; };
__device__ __forceinline__ void tile_mac(Tile& acc, const double (*As)[DB + 4], const double (*Bs)[DB + 4], int ty, int tx)
{
#pragma unroll 8
    for (int q = 0; q < DB; ++q) {
        double a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { a[u] = As[q][ty * 4 + u]; b[u] = Bs[q][tx * 4 + u]; }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc.v[u][v] += a[u] * b[v];
    }
}
__device__ __forceinline__ void tile_zero(Tile& t)
{
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) t.v[u][v] = 0.0;
}

// Y = L^{-1} (lower), one CTA per 32-wide block column j:  Y_jj = Dinv_j,
// Y_ij = -Dinv_i * sum_{k=j}^{i-1} L_ik Y_kj   (forward substitution on the identity)
__global__ void __launch_bounds__(64)
dn_trtri_cols(int n, int nb, const double* __restrict__ L, const double* __restrict__ Dinv, double* Y)
{
    __shared__ double As[DB][DB + 4], Bs[DB][DB + 4];
    const int j = blockIdx.x, t = threadIdx.x, ty = t / 8, tx = t % 8;
    for (int i = j; i < nb; ++i) {
        Tile acc;
        tile_zero(acc);
        for (int k = j; k < i; ++k) {
            for (int e = t; e < DB * DB; e += 64) {
                const int rr = e % DB, cc = e / DB;
                As[cc][rr] = (i * DB + rr < n && k * DB + cc < n) ? L[(i * DB + rr) + (size_t)(k * DB + cc) * n] : 0.0;
                Bs[rr][cc] = (k * DB + rr < n && j * DB + cc < n) ? Y[(k * DB + rr) + (size_t)(j * DB + cc) * n] : 0.0;
            }
            __syncthreads();
            tile_mac(acc, As, Bs, ty, tx);
            __syncthreads();
        }
        // S = (i == j) ? I : -acc, then Y_ij = Dinv_i * S
        for (int u = 0; u < 4; ++u)
            for (int v = 0; v < 4; ++v) {
                const int rr = ty * 4 + u, cc = tx * 4 + v;
                Bs[rr][cc] = (i == j) ? (rr == cc ? 1.0 : 0.0) : -acc.v[u][v];
            }
        for (int e = t; e < DB * DB; e += 64) { const int rr = e % DB, cc = e / DB; As[cc][rr] = Dinv[(size_t)i * DB * DB + rr + cc * DB]; }
        __syncthreads();
        Tile out;
        tile_zero(out);
        tile_mac(out, As, Bs, ty, tx);
        for (int u = 0; u < 4; ++u)
            for (int v = 0; v < 4; ++v) {
                const int r = i * DB + ty * 4 + u, c = j * DB + tx * 4 + v;
                if (r < n && c < n) Y[r + (size_t)c * n] = out.v[u][v];
            }
        __syncthreads();       // Y_ij is read back from global memory by this CTA in later steps
    }
}

// X = L^{-T} Y, lower block rows i >= j only (X = Sigma^{-1} is symmetric):
// X_ij = Dinv_i^T (Y_ij - sum_{k>i} L_ki^T X_kj), walking i from the last block row up to j
__global__ void __launch_bounds__(64)
dn_lauum_cols(int n, int nb, const double* __restrict__ L, const double* __restrict__ Dinv, const double* __restrict__ Y,
              double* X)
{
    __shared__ double As[DB][DB + 4], Bs[DB][DB + 4];
    const int j = blockIdx.x, t = threadIdx.x, ty = t / 8, tx = t % 8;
    for (int i = nb - 1; i >= j; --i) {
        Tile acc;
        tile_zero(acc);
        for (int k = i + 1; k < nb; ++k) {
            for (int e = t; e < DB * DB; e += 64) {
                const int rr = e % DB, cc = e / DB;
                // As[q][row] = (L_ki^T)[row][q] = L_ki[q][row]
                As[rr][cc] = (k * DB + rr < n && i * DB + cc < n) ? L[(k * DB + rr) + (size_t)(i * DB + cc) * n] : 0.0;
                Bs[rr][cc] = (k * DB + rr < n && j * DB + cc < n) ? X[(k * DB + rr) + (size_t)(j * DB + cc) * n] : 0.0;
            }
            __syncthreads();
            tile_mac(acc, As, Bs, ty, tx);
            __syncthreads();
        }
        for (int u = 0; u < 4; ++u)
            for (int v = 0; v < 4; ++v) {
                const int rr = ty * 4 + u, cc = tx * 4 + v;
                const int r = i * DB + rr, c = j * DB + cc;
                const double y = (r < n && c < n) ? Y[r + (size_t)c * n] : 0.0;
                Bs[rr][cc] = y - acc.v[u][v];
            }
        // As[q][row] = (Dinv_i^T)[row][q] = Dinv_i[q][row]
        for (int e = t; e < DB * DB; e += 64) { const int rr = e % DB, cc = e / DB; As[rr][cc] = Dinv[(size_t)i * DB * DB + rr + cc * DB]; }
        __syncthreads();
        Tile out;
        tile_zero(out);
        tile_mac(out, As, Bs, ty, tx);
        for (int u = 0; u < 4; ++u)
            for (int v = 0; v < 4; ++v) {
                const int r = i * DB + ty * 4 + u, c = j * DB + tx * 4 + v;
                if (r < n && c < n) X[r + (size_t)c * n] = out.v[u][v];
            }
        __syncthreads();
    }
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
    const int N = (int)n, nb = (N + DB - 1) / DB;
    const int ovw = (flags & FADB_OVERWRITE_ADJ) ? 1 : 0;
    const size_t nn = n * n;
    const int eb = (int)((nn + 255) / 256);
    cudaStream_t st = c->stream;
    if (factor) {
        FADB_CUDA(cudaMemsetAsync(w->scal, 0, 8 * sizeof(double), st));
        dn_copy_lower<<<eb, 256, 0, st>>>(N, Sigma, w->L);
        c->launches++;
        for (int k = 0; k < nb; ++k) {
            const int k0 = k * DB;
            dn_potf2<<<1, dim3(DB, DB), 0, st>>>(N, w->L, k0, w->Dinv + (size_t)k * DB * DB, w->scal);
            c->launches++;
            const int m = N - k0 - DB;
            if (m > 0) {
                dn_trsm_panel<<<(m + 63) / 64, 256, 0, st>>>(N, w->L, k0, w->Dinv + (size_t)k * DB * DB);
                const int tiles = (m + 63) / 64;
                dn_syrk<<<dim3(tiles, tiles), dim3(16, 16), 0, st>>>(N, w->L, k0);
                c->launches += 2;
            }
        }
        {                              // z is formed from the explicit inverse, as the reference does (normal.hpp:632)
            dn_trtri_cols<<<nb, 64, 0, st>>>(N, nb, w->L, w->Dinv, w->Y);
            dn_lauum_cols<<<nb, 64, 0, st>>>(N, nb, w->L, w->Dinv, w->Y, w->X);
            dn_symmetrize<<<eb, 256, 0, st>>>(N, w->X);
            c->launches += 3;
        }
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
