// dense_normal.cu — K7: normal_adj_log_pdf(x, mu, Sigma) with a dense covariance matrix
// (SURVEY.md 8f row 3): NormalAdjLogPDFNode<vec, scl|vec, mat>::feval/beval,
// reverse/stat/normal.hpp:581-773 — Eigen::LLT<Lower> of Sigma, log_det = log det L,
// inv = llt.solve(I), z = inv (x - mu), value = -0.5 (x - mu)^T z - log_det, and the adjoints
// Sigma_adj += (-0.5 seed)(inv - z z^T), mu_adj += seed z (or its sum), x_adj += -seed z.
//
// All FP64, hand-written, deterministic (no atomics):
//   1. two-level right-looking Cholesky (128-wide outer panels, 32-wide inner steps): ONE launch per
//      inner step factors the diagonal tile in registers, solves the panel against it and updates
//      the rest of the outer panel (rank 32); a tiled SYRK then updates the trailing matrix
//      (rank 128) once per outer panel;
//   2. the 32x32 diagonal tiles are inverted in one batched launch, then Y = L^{-1} by recursive halving ([A 0; B C]^-1 = [A^-1 0; -C^-1 B A^-1, C^-1]), one batched launch
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
    double *L = nullptr, *Y = nullptr, *X = nullptr, *d = nullptr, *z = nullptr, *scal = nullptr, *diag = nullptr;
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

// ---- one inner step of the blocked Cholesky, ONE launch:  diagonal tile -> panel -> rank-32 update
// of the rest of the 128-wide outer panel.  CTA b owns rows [k0+32+64b, +64) below the tile.
//  A. warp 0 of EVERY CTA factors the 32x32 diagonal tile redundantly - it is a dependent 32-step
//     chain nobody can help with, and repeating it saves two launches and a round trip through
//     global memory; the other warps meanwhile stage the raw panel rows.
//     Lane r keeps row r of the trailing tile in registers; each step updates and ROTATES the row
//     by one column (row[c-1] = row[c] - l_rj * l_(j+c)j), so the step body has static register
//     indices inside a run-time loop and the code stays small (a fully unrolled version ran at
//     instruction-fetch speed, 30 us per tile; a shared-memory one at 57 us; a 32x32-thread one
//     with block barriers at 31 us).  1/sqrt(pivot): MUFU seed + two Newton steps.
//  B. P <- P L_kk^{-T} by forward substitution, one thread per row (same rotate trick, the row
//     lives in registers), for the CTA's own 64 rows and (again redundantly) for the `pc` rows at
//     the top of the panel, whose columns the update needs.
//  C. C[own rows][top pc columns] -= P_own P_top^T, lower part only; the C values are fetched
//     before phase B so that their latency hides behind it.
// The step READS the working matrix L (diagonal tile, raw panel) and WRITES the finished factor
// columns to F: CTAs of one launch read each other's raw rows, so the panel is never updated in place.
constexpr int CS_ROWS = 64, CS_TOP = 96;
struct CholSmem {
    double Tc[DB][2 * DB];                 // Tc[j][i] = l_ij (column j contiguous), zero for i >= 32
    double colbuf[2 * DB], rinv_s[DB];
    double P[CS_ROWS + CS_TOP + 8][DB + 1];// own rows first, then the top rows of the panel (+8 rows of slack)
};

__global__ void __launch_bounds__(256)
dn_chol_step(int n, double* L, double* F, double* diag, int k0, int pc, double* flag)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CholSmem& S = *reinterpret_cast<CholSmem*>(smem_raw);
    const int t = threadIdx.x, r = t & 31;
    const int top0 = k0 + DB, own0 = top0 + blockIdx.x * CS_ROWS, R = CS_ROWS + pc;
    if (t < 32) {
        double row[DB];
#pragma unroll
        for (int c = 0; c < DB; ++c) {
            row[c] = (r >= c) ? ld_pad(L, n, k0 + r, k0 + c) : 0.0;
            S.Tc[c][DB + r] = 0.0;
        }
        S.colbuf[DB + r] = 0.0;
        bool bad = false;
        const bool writer = blockIdx.x == 0;
        for (int j = 0; j < DB; ++j) {
            const double p = __shfl_sync(0xffffffffu, row[0], j);
            bad |= !(p > 0);                                // llt.info() != Success, normal.hpp:660
            double y;
            asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(p));
            const double h = 0.5 * p;
            y = y * fma(-h * y, y, 1.5);
            y = y * fma(-h * y, y, 1.5);
            const double lrj = (r >= j) ? row[0] * y : 0.0; // l_rj; the pivot lane gets p / sqrt(p)
            S.colbuf[r] = lrj;
            S.Tc[j][r] = lrj;
            if (r == j) S.rinv_s[j] = y;
            if (writer && r >= j && k0 + r < n && k0 + j < n) F[(k0 + r) + (size_t)(k0 + j) * n] = lrj;
            if (writer && r == j && k0 + j < n) diag[k0 + j] = lrj;
            __syncwarp();
#pragma unroll
            for (int c = 1; c < DB; ++c) row[c - 1] = fma(-lrj, S.colbuf[j + c], row[c]);
            row[DB - 1] = 0.0;
            __syncwarp();
        }
        if (bad && writer && r == 0) flag[0] = 1.0;
    } else {
        for (int e = t - 32; e < R * DB; e += 224) {
            const int rr = e % R, cc = e / R;
            const int g = rr < CS_ROWS ? own0 + rr : top0 + (rr - CS_ROWS);
            S.P[rr][cc] = (g < n && k0 + cc < n) ? L[g + (size_t)(k0 + cc) * n] : 0.0;
        }
    }
    // C values of this thread (row i, columns jg*8 + 32*b + v), in flight during phase B
    const int i = t & 63, jg = t >> 6, gi = own0 + i;
    double cv[3][8];
#pragma unroll
    for (int bq = 0; bq < 3; ++bq)
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            const int j = jg * 8 + bq * 32 + v, gj = top0 + j;
            cv[bq][v] = (j < pc && gi < n && gi >= gj) ? L[gi + (size_t)gj * n] : 0.0;
        }
    __syncthreads();
    if (own0 >= n && blockIdx.x > 0) return;
    // B: forward substitution x L_kk^T = p, one row per thread, row in registers
    if (t < R) {
        double pr[DB];
#pragma unroll
        for (int q = 0; q < DB; ++q) pr[q] = S.P[t][q];
        const bool own = t < CS_ROWS && own0 + t < n;
        for (int c = 0; c < DB; ++c) {
            const double x = pr[0] * S.rinv_s[c];
            S.P[t][c] = x;
            if (own && k0 + c < n) F[(own0 + t) + (size_t)(k0 + c) * n] = x;
#pragma unroll
            for (int u = 1; u < DB; ++u) pr[u - 1] = fma(-x, S.Tc[c][c + u], pr[u]);
            pr[DB - 1] = 0.0;
        }
    }
    __syncthreads();
    // C: rank-32 update of the rest of the outer panel, 8 columns (independent chains) at a time
    if (pc <= 0) return;
    double pr[DB];
#pragma unroll
    for (int q = 0; q < DB; ++q) pr[q] = S.P[i][q];
#pragma unroll
    for (int bq = 0; bq < 3; ++bq) {
        const int jb = jg * 8 + bq * 32;
        if (jb >= pc) break;
        double acc[8] = {};
#pragma unroll
        for (int q = 0; q < DB; ++q)
#pragma unroll
            for (int v = 0; v < 8; ++v) acc[v] = fma(pr[q], S.P[CS_ROWS + jb + v][q], acc[v]);
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            const int gj = top0 + jb + v;
            if (jb + v < pc && gi < n && gi >= gj) L[gi + (size_t)gj * n] = cv[bq][v] - acc[v];
        }
    }
}

// inverses of all 32x32 diagonal tiles of the factor, in place in Y (the leaves of the recursive
// inversion), one warp per tile: column-oriented forward substitution, lane r solving column r.
__global__ void __launch_bounds__(32)
dn_tile_inverse(int n, double* Y)
{
    __shared__ double Tc[DB][2 * DB];
    const int r = threadIdx.x, k0 = blockIdx.x * DB;
#pragma unroll
    for (int c = 0; c < DB; ++c) {
        Tc[c][r] = (r >= c) ? ld_pad(Y, n, k0 + r, k0 + c) : 0.0;
        Tc[c][DB + r] = 0.0;
    }
    __syncwarp();
    double tt[DB];
#pragma unroll
    for (int u = 0; u < DB; ++u) tt[u] = (u == r) ? 1.0 : 0.0;
    for (int i = 0; i < DB; ++i) {
        const double v = tt[0] / Tc[i][i];                  // (L_kk^-1)[i][r]
        if (k0 + i < n && k0 + r < n) Y[(k0 + i) + (size_t)(k0 + r) * n] = v;
#pragma unroll
        for (int u = 1; u < DB; ++u) tt[u - 1] = fma(-Tc[i][i + u], v, tt[u]);
        tt[DB - 1] = 0.0;
    }
}

// ---- general FP64 GEMM: C = alpha * op(A) * op(B) + beta * C, column-major.  256 threads own a
// BT x BT tile (BT = 64: 4x4 outputs per thread, K step 16; BT = 128: 8x8 outputs, K step 8).
//  * the next K slab is fetched from global memory into registers while the current one is
//    multiplied out of shared memory (two shared buffers, one barrier per slab);
//  * a warp covers 8 x 4 threads and every thread reads its fragments as 16-byte pairs at
//    (chunk * 32 + 2 * thread) so that the LDS.128s of a warp are conflict-free and 8 lanes
//    broadcast; rows run along the fast thread index so stores to C coalesce.
// Triangular structure is exploited at tile granularity through `kmode`:
//   0 full K;  1 op(B) is lower triangular (K starts at the tile's first column);
//   2 op(A) is lower triangular (K ends at the tile's last row);
//   3 op(A) = M^T, op(B) = M with M lower triangular (K starts at max(tile row, tile col)).
// lower_only = 1 computes the tiles on or below the diagonal and masks the strict upper part.
template <int BT, bool TA, bool TB>
__device__ __forceinline__ void gemm_tile(int bi, int bj, int m, int n, int k, double alpha, const double* __restrict__ A,
                                          int lda, const double* __restrict__ B, int ldb, double beta, double* C, int ldc,
                                          int kmode, int lower_only)
{
    constexpr int BK = BT == 128 ? 8 : 16, TM = BT / 16, SS = BT + 2, NL = BT * BK / 256;
    __shared__ __align__(16) double As[2][BK][SS], Bs[2][BK][SS];
    if (lower_only && bj > bi) return;
    const int i0 = bi * BT, j0 = bj * BT;
    if (i0 >= m || j0 >= n) return;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int tx = (w & 1) * 8 + (lane & 7), ty = (w >> 1) * 4 + (lane >> 3);
    int kb = 0, ke = k;
    if (kmode == 1) kb = j0;
    else if (kmode == 2) ke = min(k, i0 + BT);
    else if (kmode == 3) kb = max(i0, j0);
    kb &= ~(BK - 1);
    double acc[TM][TM] = {};
    double ra[NL], rb[NL];
    auto gload = [&](int k0) {
#pragma unroll
        for (int u = 0; u < NL; ++u) {
            const int e = t + u * 256;
            const int ii = TA ? e / BK : e % BT, kk = TA ? e % BK : e / BT;
            const int gi = i0 + ii, gk = k0 + kk;
            ra[u] = (gi < m && gk < ke) ? (TA ? A[gk + (size_t)gi * lda] : A[gi + (size_t)gk * lda]) : 0.0;
            const int jj = TB ? e % BT : e / BK, kq = TB ? e / BT : e % BK;
            const int gj = j0 + jj, gq = k0 + kq;
            rb[u] = (gj < n && gq < ke) ? (TB ? B[gj + (size_t)gq * ldb] : B[gq + (size_t)gj * ldb]) : 0.0;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int u = 0; u < NL; ++u) {
            const int e = t + u * 256;
            As[buf][TA ? e % BK : e / BT][TA ? e / BK : e % BT] = ra[u];
            Bs[buf][TB ? e / BT : e % BK][TB ? e % BT : e / BK] = rb[u];
        }
    };
    gload(kb);
    sstore(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = kb; k0 < ke; k0 += BK) {
        const bool more = k0 + BK < ke;
        if (more) gload(k0 + BK);
#pragma unroll
        for (int q = 0; q < BK; ++q) {
            double a[TM], b[TM];
#pragma unroll
            for (int c = 0; c < TM / 2; ++c) {
                const double2 va = *reinterpret_cast<const double2*>(&As[buf][q][c * 32 + tx * 2]);
                const double2 vb = *reinterpret_cast<const double2*>(&Bs[buf][q][c * 32 + ty * 2]);
                a[2 * c] = va.x; a[2 * c + 1] = va.y;
                b[2 * c] = vb.x; b[2 * c + 1] = vb.y;
            }
#pragma unroll
            for (int u = 0; u < TM; ++u)
#pragma unroll
                for (int v = 0; v < TM; ++v) acc[u][v] = fma(a[u], b[v], acc[u][v]);
        }
        if (more) {
            sstore(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }
#pragma unroll
    for (int v = 0; v < TM; ++v)
#pragma unroll
        for (int u = 0; u < TM; ++u) {
            const int i = i0 + (u / 2) * 32 + tx * 2 + (u & 1), j = j0 + (v / 2) * 32 + ty * 2 + (v & 1);
            if (i < m && j < n && (!lower_only || i >= j)) {
                double* c = C + i + (size_t)j * ldc;
                *c = (beta == 0.0) ? alpha * acc[u][v] : fma(alpha, acc[u][v], beta * *c);
            }
        }
}

template <int BT, bool TA, bool TB>
__global__ void __launch_bounds__(256, BT == 128 ? 1 : 3)
dn_gemm(int m, int n, int k, double alpha, const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
        double beta, double* C, int ldc, int kmode, int lower_only)
{
    gemm_tile<BT, TA, TB>(blockIdx.x, blockIdx.y, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, kmode, lower_only);
}

// ---- batched triangular inversion.  Node z of the recursion tree owns W[off : off+n1+n2]^2 =
// [A 0; B C] with A^-1 (n1 x n1) and C^-1 (n2 x n2) already in place; all nodes of equal height are
// independent, so one launch handles a whole level:  phase 0: T_z = B A^-1,  phase 1: B <- -C^-1 T_z.
struct TrNode { int off, n1, n2, pad; long long toff; };

template <int BT>
__global__ void __launch_bounds__(256, BT == 128 ? 1 : 3)
dn_trtri_level(int n, double* W, double* T, const TrNode* __restrict__ nodes, int phase)
{
    const TrNode nd = nodes[blockIdx.z];
    double* Ainv = W + nd.off + (size_t)nd.off * n;
    double* Bm = W + (nd.off + nd.n1) + (size_t)nd.off * n;
    double* Cinv = W + (nd.off + nd.n1) + (size_t)(nd.off + nd.n1) * n;
    double* Tz = T + nd.toff;
    if (phase == 0)
        gemm_tile<BT, false, false>(blockIdx.x, blockIdx.y, nd.n2, nd.n1, nd.n1, 1.0, Bm, n, Ainv, n, 0.0, Tz, nd.n2, 1, 0);
    else
        gemm_tile<BT, false, false>(blockIdx.x, blockIdx.y, nd.n2, nd.n1, nd.n2, -1.0, Cinv, n, Tz, nd.n2, 0.0, Bm, n, 2, 0);
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
dn_finish(int n, const double* __restrict__ diag, const double* __restrict__ d, const double* __restrict__ z,
          const double* flag, double seed, int overwrite, double* x_adj, double* mu_adj, int mu_vec, double* out_value)
{
    __shared__ double sm[32 * 3];
    double acc[3] = {0.0, 0.0, 0.0};             // d^T z, sum log L_ii, sum z
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        acc[0] += d[i] * z[i];
        acc[1] += log(diag[i]);
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

// 128-wide tiles (8x8 outputs per thread) once they fill the GPU, 64-wide ones below that
static inline bool big_tiles(const Context& c, long long tiles128) { return tiles128 >= (long long)c.sms; }

template <bool TA, bool TB>
static void gemm(Context& c, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb,
                 double beta, double* C, int ldc, int kmode, int lower_only)
{
    if (m <= 0 || n <= 0) return;
    const long long t128 = (long long)((m + 127) / 128) * ((n + 127) / 128) / (lower_only ? 2 : 1);
    if (big_tiles(c, t128)) {
        dim3 grid((m + 127) / 128, (n + 127) / 128);
        dn_gemm<128, TA, TB><<<grid, 256, 0, c.stream>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, kmode, lower_only);
    } else {
        dim3 grid((m + 63) / 64, (n + 63) / 64);
        dn_gemm<64, TA, TB><<<grid, 256, 0, c.stream>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, kmode, lower_only);
    }
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
        for (double* p : {w.L, w.Y, w.X, w.d, w.z, w.scal, w.diag}) if (p) FADB_CUDA(cudaFree(p));
        FADB_CUDA(cudaMalloc(&w.L, n * n * sizeof(double)));
        FADB_CUDA(cudaMalloc(&w.Y, n * n * sizeof(double)));
        FADB_CUDA(cudaMalloc(&w.X, n * n * sizeof(double)));
        FADB_CUDA(cudaMalloc(&w.d, n * sizeof(double)));
        FADB_CUDA(cudaMalloc(&w.z, n * sizeof(double)));
        FADB_CUDA(cudaMalloc(&w.diag, n * sizeof(double)));
        FADB_CUDA(cudaMalloc(&w.scal, 8 * sizeof(double)));
        FADB_CUDA(cudaFuncSetAttribute(dn_chol_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CholSmem)));
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
        FADB_CUDA(cudaMemsetAsync(w->Y, 0, nn * sizeof(double), st));      // the factor is written into Y, lower part only
        // two-level right-looking Cholesky: 128-wide outer panels, 32-wide inner steps.  The inner
        // rank-32 update touches the rest of the outer panel only; the trailing matrix gets one
        // rank-128 update per outer panel.
        constexpr int OB = 128;
        for (int K0 = 0; K0 < N; K0 += OB) {
            const int W = std::min(OB, N - K0), Kend = K0 + W;
            for (int k0 = K0; k0 < Kend; k0 += DB) {
                const int m = N - k0 - DB, pc = Kend - (k0 + DB);
                const int grid = m > 0 ? (m + CS_ROWS - 1) / CS_ROWS : 1;
                dn_chol_step<<<grid, 256, sizeof(CholSmem), st>>>(N, w->L, w->Y, w->diag, k0, m > 0 ? pc : 0, w->scal);
                c->launches++;
            }
            const int m = N - Kend;
            if (m > 0) {
                const double* P = w->Y + Kend + (size_t)K0 * N;
                double* Ct = w->L + Kend + (size_t)Kend * N;
                gemm<false, true>(*c, m, m, W, -1.0, P, N, P, N, 1.0, Ct, N, 0, 1);      // trailing -= P P^T (lower)
            }
        }
        // inverse, in place: Y <- Y^{-1} level by level up the halving tree, X <- Y^T Y (lower tiles, K from the diagonal)
        dn_tile_inverse<<<nb, 32, 0, st>>>(N, w->Y);
        c->launches++;
        for (const TrLevel& lv : tp->levels) {
            const long long t128 = (long long)((lv.max_n2 + 127) / 128) * ((lv.max_n1 + 127) / 128) * lv.count;
            if (big_tiles(*c, t128)) {
                dim3 grid((lv.max_n2 + 127) / 128, (lv.max_n1 + 127) / 128, lv.count);
                dn_trtri_level<128><<<grid, 256, 0, st>>>(N, w->Y, w->X, tp->dev + lv.first, 0);
                dn_trtri_level<128><<<grid, 256, 0, st>>>(N, w->Y, w->X, tp->dev + lv.first, 1);
            } else {
                dim3 grid((lv.max_n2 + 63) / 64, (lv.max_n1 + 63) / 64, lv.count);
                dn_trtri_level<64><<<grid, 256, 0, st>>>(N, w->Y, w->X, tp->dev + lv.first, 0);
                dn_trtri_level<64><<<grid, 256, 0, st>>>(N, w->Y, w->X, tp->dev + lv.first, 1);
            }
            c->launches += 2;
        }
        gemm<true, false>(*c, N, N, N, 1.0, w->Y, N, w->Y, N, 0.0, w->X, N, 3, 1);
        dn_symmetrize<<<eb, 256, 0, st>>>(N, w->X);
        c->launches++;
    }
    dn_gemv<<<(N * 32 + 255) / 256, 256, 0, st>>>(N, w->X, x, mu, mu_vec, w->d, w->z);
    dn_finish<<<1, 1024, 0, st>>>(N, w->diag, w->d, w->z, w->scal, seed, ovw, x_adj, mu_adj, mu_vec, dev_value);
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
