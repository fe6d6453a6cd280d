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
    int ready;                             // columns of the tile finished by warp 0 (phase B follows it column by column)
};

__global__ void __launch_bounds__(256)
dn_chol_step(int n, double* L, double* F, double* diag, int k0, int pc, double* flag)
{
    pdl_prologue();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CholSmem& S = *reinterpret_cast<CholSmem*>(smem_raw);
    const int t = threadIdx.x, r = t & 31;
    const int top0 = k0 + DB, own0 = top0 + blockIdx.x * CS_ROWS, R = CS_ROWS + pc;
    if (t == 0) S.ready = 0;
    __syncthreads();
    // C values of this thread (row i, columns jg*8 + 32*b + v), in flight during phases A and B
    const int i = t & 63, jg = t >> 6, gi = own0 + i;
    double cv[3][8];
#pragma unroll
    for (int bq = 0; bq < 3; ++bq)
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            const int j = jg * 8 + bq * 32 + v, gj = top0 + j;
            cv[bq][v] = (j < pc && gi < n && gi >= gj) ? L[gi + (size_t)gj * n] : 0.0;
        }
    // %globaltimer stamps of one step at n = 4000 (round 2): 8 us for the 32-pivot chain of phase A, THEN 4 us for the forward
    // substitution of phase B, one thread per row.  B needs column j of the tile, not the tile: the row threads (warps 1..)
    // now follow warp 0 column by column (S.ready) and finish a few hundred cycles after it.
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
        double p = __shfl_sync(0xffffffffu, row[0], 0);
        for (int j = 0; j < DB; ++j) {
            bad |= !(p > 0);                                // llt.info() != Success, normal.hpp:660
            double y;
            asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(p));
            const double h = 0.5 * p;
            y = y * fma(-h * y, y, 1.5);
            y = y * fma(-h * y, y, 1.5);
            const double lrj = (r >= j) ? row[0] * y : 0.0; // l_rj; the pivot lane gets p / sqrt(p)
            // the NEXT pivot is lane j+1's a_(j+1)(j+1) - l_(j+1)j^2: formed from registers, ahead of the update sweep (the
            // sweep computes the same fma through shared memory), so that its rsqrt chain starts without the round trip
            const double own_next = fma(-lrj, lrj, row[1]);
            p = __shfl_sync(0xffffffffu, own_next, (j + 1) & 31);
            S.colbuf[r] = lrj;
            S.Tc[j][r] = lrj;
            if (r == j) S.rinv_s[j] = y;
            if (writer && r >= j && k0 + r < n && k0 + j < n) F[(k0 + r) + (size_t)(k0 + j) * n] = lrj;
            if (writer && r == j && k0 + j < n) diag[k0 + j] = lrj;
            __syncwarp();
            __threadfence_block();
            if (r == 0) *reinterpret_cast<volatile int*>(&S.ready) = j + 1;
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
        asm volatile("bar.sync 1, 224;" ::: "memory");      // the raw rows are staged (warps 1..7 only; warp 0 is factoring)
        // B: forward substitution x L_kk^T = p, one row per thread, row in registers, one column behind warp 0
        const int tt = t - 32;
        if (tt < R) {
            double pr[DB];
#pragma unroll
            for (int q = 0; q < DB; ++q) pr[q] = S.P[tt][q];
            const bool own = tt < CS_ROWS && own0 + tt < n;
            for (int c = 0; c < DB; ++c) {
                while (*reinterpret_cast<volatile int*>(&S.ready) <= c) __nanosleep(20);
                __threadfence_block();
                const double x = pr[0] * S.rinv_s[c];
                S.P[tt][c] = x;
                if (own && k0 + c < n) F[(own0 + tt) + (size_t)(k0 + c) * n] = x;
#pragma unroll
                for (int u = 1; u < DB; ++u) pr[u - 1] = fma(-x, S.Tc[c][c + u], pr[u]);
                pr[DB - 1] = 0.0;
            }
        }
    }
    __syncthreads();
    if (own0 >= n && blockIdx.x > 0) return;
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
    pdl_prologue();
    gemm_tile<BT, TA, TB>(blockIdx.x, blockIdx.y, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, kmode, lower_only);
}

// ---- the same tile on the FP64 TENSOR cores (BASELINE.json north_star: "FP64 DMMA ... only if ncu shows they pay").
// mma.sync.aligned.m8n8k4.f64 (SASS DMMA.8x8x4): one warp instruction does 8 x 8 x 4 = 256 FMAs against 32 for a DFMA, with
// A / B fragments of ONE double per thread -- the register-tiled kernel above spends one LDS.128 per 4 FMAs on operand
// traffic and 2 operand registers per FMA, this one one LDS.64 per 21 FMAs.  Same staging (global -> registers -> shared,
// double-buffered), same kmode / lower_only tile skipping, same epilogue; only the summation order inside a tile differs.
// 8 warps as WM x WN, each owning MT x NT fragments of 8 x 8.  Shared rows are padded to SS = BT + 4 doubles: the fragment
// loads As[k + lane % 4][m + lane / 4] of a half-warp then fall on 16 distinct bank pairs (row stride = 8 banks mod 32).
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int BT, bool TA, bool TB>
__device__ __forceinline__ void gemm_tile_mma(int bi, int bj, int m, int n, int k, double alpha, const double* __restrict__ A,
                                              int lda, const double* __restrict__ B, int ldb, double beta, double* C, int ldc,
                                              int kmode, int lower_only)
{
    constexpr int BK = BT == 128 ? 8 : 16, SS = BT + 4, NL = BT * BK / 256;
    constexpr int WM = BT == 128 ? 4 : 2, WN = 8 / WM, MT = BT / WM / 8, NT = BT / WN / 8;
    __shared__ __align__(16) double As[2][BK][SS], Bs[2][BK][SS];
    if (lower_only && bj > bi) return;
    const int i0 = bi * BT, j0 = bj * BT;
    if (i0 >= m || j0 >= n) return;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int wm = w % WM, wn = w / WM, g = lane >> 2, q4 = lane & 3;
    int kb = 0, ke = k;
    if (kmode == 1) kb = j0;
    else if (kmode == 2) ke = min(k, i0 + BT);
    else if (kmode == 3) kb = max(i0, j0);
    kb &= ~(BK - 1);
    double acc[MT][NT][2] = {};
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
        for (int q = 0; q < BK; q += 4) {
            double a[MT], b[NT];
#pragma unroll
            for (int u = 0; u < MT; ++u) a[u] = As[buf][q + q4][wm * (MT * 8) + u * 8 + g];
#pragma unroll
            for (int v = 0; v < NT; ++v) b[v] = Bs[buf][q + q4][wn * (NT * 8) + v * 8 + g];
#pragma unroll
            for (int u = 0; u < MT; ++u)
#pragma unroll
                for (int v = 0; v < NT; ++v) dmma884(acc[u][v][0], acc[u][v][1], a[u], b[v]);
        }
        if (more) {
            sstore(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }
#pragma unroll
    for (int v = 0; v < NT; ++v)
#pragma unroll
        for (int u = 0; u < MT; ++u)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = i0 + wm * (MT * 8) + u * 8 + g, j = j0 + wn * (NT * 8) + v * 8 + q4 * 2 + h;
                if (i < m && j < n && (!lower_only || i >= j)) {
                    double* c = C + i + (size_t)j * ldc;
                    *c = (beta == 0.0) ? alpha * acc[u][v][h] : fma(alpha, acc[u][v][h], beta * *c);
                }
            }
}

template <int BT, bool TA, bool TB>
__global__ void __launch_bounds__(256, BT == 128 ? 1 : 3)
dn_gemm_mma(int m, int n, int k, double alpha, const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
            double beta, double* C, int ldc, int kmode, int lower_only)
{
    pdl_prologue();
    int bi = blockIdx.x, bj = blockIdx.y;
    if (lower_only && gridDim.y == 1) {
        // triangular launch (square products only): CTA b is tile (bi, bj), bj <= bi, b = bi (bi + 1) / 2 + bj -- the upper half
        // of a square grid is CTAs that start only to return (1800 of 3600 per rank-128 update at n = 4000)
        bi = (int)((sqrt(8.0 * blockIdx.x + 1.0) - 1.0) * 0.5);
        while ((bi + 1) * (bi + 2) / 2 <= (int)blockIdx.x) ++bi;
        while (bi * (bi + 1) / 2 > (int)blockIdx.x) --bi;
        bj = (int)blockIdx.x - bi * (bi + 1) / 2;
    }
    gemm_tile_mma<BT, TA, TB>(bi, bj, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, kmode, lower_only);
}

// FADB_GEMM_DMMA=0 selects the DFMA register-tiled kernels everywhere (A/B runs); default: tensor cores
static inline bool use_dmma()
{
    static const bool on = [] { const char* e = getenv("FADB_GEMM_DMMA"); return !(e && e[0] == '0'); }();
    return on;
}

// ---- batched triangular inversion.  Node z of the recursion tree owns W[off : off+n1+n2]^2 =
// [A 0; B C] with A^-1 (n1 x n1) and C^-1 (n2 x n2) already in place; all nodes of equal height are
// independent, so one launch handles a whole level:  phase 0: T_z = B A^-1,  phase 1: B <- -C^-1 T_z.
struct TrNode { int off, n1, n2, pad; long long toff; };

template <int BT>
__global__ void __launch_bounds__(256, BT == 128 ? 1 : 3)
dn_trtri_level(int n, double* W, double* T, const TrNode* __restrict__ nodes, int phase)
{
    pdl_prologue();
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
template <int BT>
__global__ void __launch_bounds__(256, BT == 128 ? 1 : 3)
dn_trtri_level_mma(int n, double* W, double* T, const TrNode* __restrict__ nodes, int phase)
{
    pdl_prologue();
    const TrNode nd = nodes[blockIdx.z];
    double* Ainv = W + nd.off + (size_t)nd.off * n;
    double* Bm = W + (nd.off + nd.n1) + (size_t)nd.off * n;
    double* Cinv = W + (nd.off + nd.n1) + (size_t)(nd.off + nd.n1) * n;
    double* Tz = T + nd.toff;
    if (phase == 0)
        gemm_tile_mma<BT, false, false>(blockIdx.x, blockIdx.y, nd.n2, nd.n1, nd.n1, 1.0, Bm, n, Ainv, n, 0.0, Tz, nd.n2, 1, 0);
    else
        gemm_tile_mma<BT, false, false>(blockIdx.x, blockIdx.y, nd.n2, nd.n1, nd.n2, -1.0, Cinv, n, Tz, nd.n2, 0.0, Bm, n, 2, 0);
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

// Tile size.  Round 1 took 128-wide tiles (8x8 outputs per thread, one CTA per SM) once they filled the GPU.  Phase timing in
// round 2 (FADB_DENSE_TIMING=1, tools/dense_probe.py) says the 64-wide tile (3 CTAs per SM) is never slower and much faster when
// K is short: at n = 4000 the rank-128 trailing updates of the Cholesky ran at 10 TFLOP/s on 128-wide tiles and at 15.7 on 64-wide
// ones, the triangular-inversion levels 1.56 / 1.34 (mixed) / 1.20 ms, Y^T Y (K = n) 0.88 / 0.87 ms: 6.15 -> 5.30 ms for the whole
// evaluation.  64 is the default; FADB_GEMM_TILE=128 restores the old rule for A/B runs.
static inline bool big_tiles(const Context& c, long long tiles128, int k)
{
    static const int forced = getenv("FADB_GEMM_TILE") ? atoi(getenv("FADB_GEMM_TILE")) : 0;
    (void)k;
    return forced == 128 && tiles128 >= (long long)c.sms;
}

template <bool TA, bool TB>
static void gemm(Context& c, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb,
                 double beta, double* C, int ldc, int kmode, int lower_only)
{
    if (m <= 0 || n <= 0) return;
    const long long t128 = (long long)((m + 127) / 128) * ((n + 127) / 128) / (lower_only ? 2 : 1);
    if (big_tiles(c, t128, k)) {
        dim3 grid((m + 127) / 128, (n + 127) / 128);
        if (use_dmma()) launch_pdl(dn_gemm_mma<128, TA, TB>, grid, dim3(256), 0, c.stream, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, kmode, lower_only);
        else launch_pdl(dn_gemm<128, TA, TB>, grid, dim3(256), 0, c.stream, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, kmode, lower_only);
    } else {
        dim3 grid((m + 63) / 64, (n + 63) / 64);
        static const bool tri = [] { const char* e = getenv("FADB_GEMM_TRI"); return !(e && e[0] == '0'); }();       // =0: A/B runs
        if (tri && lower_only && m == n && use_dmma() && grid.x > 1) grid = dim3(grid.x * (grid.x + 1) / 2, 1);
        if (use_dmma()) launch_pdl(dn_gemm_mma<64, TA, TB>, grid, dim3(256), 0, c.stream, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, kmode, lower_only);
        else launch_pdl(dn_gemm<64, TA, TB>, grid, dim3(256), 0, c.stream, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, kmode, lower_only);
    }
    c.launches++;
}

// C = alpha * op(A) * op(B) + beta * C for other translation units (DotNode with large matrix operands, expr.cu)
void gemm_f64(Context& c, bool ta, bool tb, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb,
              double beta, double* C, int ldc)
{
    if (!ta && !tb) gemm<false, false>(c, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, 0, 0);
    else if (ta && !tb) gemm<true, false>(c, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, 0, 0);
    else if (!ta && tb) gemm<false, true>(c, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, 0, 0);
    else gemm<true, true>(c, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, 0, 0);
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

// Cholesky of the LOWER triangle of Sigma, then its explicit inverse: on return w->diag = L_ii,
// w->X = Sigma^{-1} (full, symmetric), w->scal[0] != 0 if Sigma is not positive definite.
// FADB_DENSE_TIMING=1: device time of the phases of one factorisation + inversion, printed to stderr (diagnostic)
struct PhaseTimer {
    bool on;
    cudaStream_t st;
    std::vector<cudaEvent_t> ev;
    std::vector<const char*> names;
    PhaseTimer(cudaStream_t s) : st(s) { static const bool e = getenv("FADB_DENSE_TIMING") != nullptr; on = e; mark("start"); }
    void mark(const char* name)
    {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        ev.push_back(e);
        names.push_back(name);
    }
    ~PhaseTimer()
    {
        if (!on) return;
        cudaEventSynchronize(ev.back());
        fprintf(stderr, "[fadb dense]");
        for (size_t k = 1; k < ev.size(); ++k) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev[k - 1], ev[k]);
            fprintf(stderr, " %s %.3f ms", names[k], ms);
        }
        fprintf(stderr, "\n");
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
    }
};

static int spd_factor_inverse(Context* c, DenseWs* w, TrPlan* tp, int N, const double* Sigma)
{
    PhaseTimer pt(c->stream);
    const int nb = (N + DB - 1) / DB;
    const size_t nn = (size_t)N * N;
    const int eb = (int)((nn + 255) / 256);
    cudaStream_t st = c->stream;
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
            launch_pdl(dn_chol_step, dim3(grid), dim3(256), sizeof(CholSmem), st, N, w->L, w->Y, w->diag, k0, m > 0 ? pc : 0, w->scal);
            c->launches++;
        }
        const int m = N - Kend;
        if (m > 0) {
            const double* P = w->Y + Kend + (size_t)K0 * N;
            double* Ct = w->L + Kend + (size_t)Kend * N;
            if (pt.on) pt.mark("steps");
            gemm<false, true>(*c, m, m, W, -1.0, P, N, P, N, 1.0, Ct, N, 0, 1);      // trailing -= P P^T (lower)
            if (pt.on) pt.mark("syrk");
        }
    }
    pt.mark("cholesky(steps+syrk)");
    // inverse, in place: Y <- Y^{-1} level by level up the halving tree, X <- Y^T Y (lower tiles, K from the diagonal)
    dn_tile_inverse<<<nb, 32, 0, st>>>(N, w->Y);
    c->launches++;
    for (const TrLevel& lv : tp->levels) {
        const long long t128 = (long long)((lv.max_n2 + 127) / 128) * ((lv.max_n1 + 127) / 128) * lv.count;
        if (big_tiles(*c, t128, lv.max_n1)) {
            dim3 grid((lv.max_n2 + 127) / 128, (lv.max_n1 + 127) / 128, lv.count);
            if (use_dmma()) {
                launch_pdl(dn_trtri_level_mma<128>, grid, dim3(256), 0, st, N, w->Y, w->X, tp->dev + lv.first, 0);
                launch_pdl(dn_trtri_level_mma<128>, grid, dim3(256), 0, st, N, w->Y, w->X, tp->dev + lv.first, 1);
            } else {
                launch_pdl(dn_trtri_level<128>, grid, dim3(256), 0, st, N, w->Y, w->X, tp->dev + lv.first, 0);
                launch_pdl(dn_trtri_level<128>, grid, dim3(256), 0, st, N, w->Y, w->X, tp->dev + lv.first, 1);
            }
        } else {
            dim3 grid((lv.max_n2 + 63) / 64, (lv.max_n1 + 63) / 64, lv.count);
            if (use_dmma()) {
                launch_pdl(dn_trtri_level_mma<64>, grid, dim3(256), 0, st, N, w->Y, w->X, tp->dev + lv.first, 0);
                launch_pdl(dn_trtri_level_mma<64>, grid, dim3(256), 0, st, N, w->Y, w->X, tp->dev + lv.first, 1);
            } else {
                launch_pdl(dn_trtri_level<64>, grid, dim3(256), 0, st, N, w->Y, w->X, tp->dev + lv.first, 0);
                launch_pdl(dn_trtri_level<64>, grid, dim3(256), 0, st, N, w->Y, w->X, tp->dev + lv.first, 1);
            }
        }
        c->launches += 2;
    }
    pt.mark("trtri");
    gemm<true, false>(*c, N, N, N, 1.0, w->Y, N, w->Y, N, 0.0, w->X, N, 3, 1);
    pt.mark("YtY");
    dn_symmetrize<<<eb, 256, 0, st>>>(N, w->X);
    c->launches++;
    pt.mark("symmetrize");
    return FADB_OK;
}


// ---------------------------------------------------------------------------------------------
// det / log_det (reverse/core/det.hpp:25-92, log_det.hpp) and wishart_adj_log_pdf
// (reverse/stat/wishart.hpp:92-231) on the same machinery.
//  * DetLLT / LogDetLLT (det.hpp:170-202): the Cholesky pipeline above; det = (prod L_ii)^2,
//    log_det = 2 sum log L_ii (the reference's log(prod) overflows for large n), inverse = X.
//  * DetFullPivLU (the default) and DetLDLT (det.hpp:97-166): a general matrix.  Determinant and
//    inverse come from an in-place Gauss-Jordan elimination with PARTIAL pivoting, two launches
//    per column (pivot search + row swap + scaling by one CTA, then a rank-1 update across the
//    GPU).  Complete pivoting / symmetric pivoting would serialise a full-matrix search per step;
//    the value and inverse agree with Eigen's to rounding, and the validity rules of the policies
//    (FullPivLU::isInvertible's epsilon * n * max-pivot threshold; LDLT: det != 0) are applied to
//    these pivots.  This path is simple rather than fast (2n launches).
// stats[0] invalid flag, [1] prod of pivots (signed), [2] sum log|pivot|, [3] min|pivot|, [4] max|pivot|
__device__ __forceinline__ void gj_pivot_body(int n, double* W, int k, double* rowbuf, double* colbuf, int* perm, double* stats)
{
    __shared__ double sval[32];
    __shared__ int sidx[32];
    __shared__ int s_p;
    __shared__ double s_piv;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    double best = -1.0;
    int bi = k;
    for (int i = k + t; i < n; i += blockDim.x) {
        const double a = fabs(W[i + (size_t)k * n]);
        if (a > best) { best = a; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_down_sync(0xffffffffu, best, o);
        const int oi = __shfl_down_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) { sval[wid] = best; sidx[wid] = bi; }
    __syncthreads();
    if (t == 0) {
        for (int w2 = 1; w2 < (int)(blockDim.x >> 5); ++w2)
            if (sval[w2] > best || (sval[w2] == best && sidx[w2] < bi)) { best = sval[w2]; bi = sidx[w2]; }
        if (!(best >= 0)) bi = k;                        // NaN column: keep the row, the flag is raised below
        const double piv = W[bi + (size_t)k * n];
        s_p = bi; s_piv = piv;
        perm[k] = bi;
        const double ap = fabs(piv);
        if (k == 0) { stats[1] = 1.0; stats[2] = 0.0; stats[3] = ap; stats[4] = ap; }
        stats[1] *= (bi != k) ? -piv : piv;
        stats[2] += log(ap);
        stats[3] = fmin(stats[3], ap);
        stats[4] = fmax(stats[4], ap);
        if (!(ap > 0)) stats[0] = 1.0;
    }
    __syncthreads();
    const int p = s_p;
    const double rinv = 1.0 / s_piv;
    for (int j = t; j < n; j += blockDim.x) {            // swap rows k and p, scale the pivot row
        const double a = W[p + (size_t)j * n], b = W[k + (size_t)j * n];
        if (p != k) W[p + (size_t)j * n] = b;
        const double r = (j == k) ? rinv : a * rinv;
        W[k + (size_t)j * n] = r;
        rowbuf[j] = r;
    }
    __syncthreads();
    for (int i = t; i < n; i += blockDim.x) {            // multipliers; column k of the inverse-in-progress
        if (i == k) { colbuf[i] = 0.0; continue; }
        const double f = W[i + (size_t)k * n];
        colbuf[i] = f;
        W[i + (size_t)k * n] = -f * rinv;
    }
}
__global__ void __launch_bounds__(1024)
gj_pivot_kernel(int n, double* W, int k, double* rowbuf, double* colbuf, int* perm, double* stats)
{
    gj_pivot_body(n, W, k, rowbuf, colbuf, perm, stats);
}

__global__ void gj_update_kernel(int n, double* W, int k, const double* __restrict__ rowbuf, const double* __restrict__ colbuf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double f = colbuf[i];
    const int j0 = blockIdx.y * 16;
#pragma unroll 4
    for (int j = j0; j < min(n, j0 + 16); ++j)
        if (j != k) W[i + (size_t)j * n] = fma(-f, rowbuf[j], W[i + (size_t)j * n]);
}

// the row swaps computed (P A)^-1 = A^-1 P^-1: undo them on the columns, last swap first
__global__ void __launch_bounds__(1024)
gj_colmap_kernel(int n, const int* __restrict__ perm, int* idx)
{
    extern __shared__ int sidx_dyn[];
    for (int j = threadIdx.x; j < n; j += blockDim.x) sidx_dyn[j] = j;
    __syncthreads();
    if (threadIdx.x == 0)
        for (int k = n - 1; k >= 0; --k) {
            const int p = perm[k];
            const int a = sidx_dyn[k]; sidx_dyn[k] = sidx_dyn[p]; sidx_dyn[p] = a;
        }
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) idx[j] = sidx_dyn[j];
}
__global__ void gj_gather_kernel(int n, const double* __restrict__ W, const int* __restrict__ idx, double* Out)
{
    const size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (e >= (size_t)n * n) return;
    const int i = (int)(e % n), j = (int)(e / n);
    Out[e] = W[i + (size_t)idx[j] * n];
}

// Small matrices (n <= GJ_SMALL_MAX, what det / log_det see in statistical models): the WHOLE elimination, the column map and
// the gather in ONE launch of one 1024-thread CTA with the matrix in SHARED memory (n^2 doubles <= 200 KB) -- 2n + 2 launches of
// mostly idle grids otherwise.  Measured (tools/det_probe.py, log_det value + adjoint): n = 4 0.094 -> 0.06 ms, n = 64 0.97 ->
// 0.3 ms.  With the matrix left in L2 a single CTA loses beyond n ~ 200 (n = 300: 9.0 against 4.5 ms), hence the shared-memory
// bound.  Same pivot rule, same fma per element and the same order of the pivot statistics as the multi-launch path (it calls
// the same gj_pivot_body), so results are bit-identical.
constexpr int GJ_SMALL_MAX = 160;
__global__ void __launch_bounds__(1024)
gj_small_kernel(int n, const double* __restrict__ A, double* rowbuf, double* colbuf, int* perm, double* stats, int* idx, double* Out)
{
    extern __shared__ __align__(16) double s_W[];        // n x n, column-major
    __shared__ double s_row[GJ_SMALL_MAX], s_col[GJ_SMALL_MAX];
    __shared__ int s_idx[GJ_SMALL_MAX];
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5, nw = blockDim.x >> 5;
    for (int e = t; e < n * n; e += blockDim.x) s_W[e] = A[e];
    __syncthreads();
    for (int k = 0; k < n; ++k) {
        gj_pivot_body(n, s_W, k, s_row, s_col, perm, stats);
        __syncthreads();
        for (int j = wid; j < n; j += nw) {              // rank-1 update, one column per warp, rows along the lanes
            if (j == k) continue;
            const double rj = s_row[j];
            double* col = s_W + (size_t)j * n;
            for (int i = lane; i < n; i += 32) col[i] = fma(-s_col[i], rj, col[i]);
        }
        __syncthreads();
    }
    (void)rowbuf; (void)colbuf;
    for (int j = t; j < n; j += blockDim.x) s_idx[j] = j;
    __syncthreads();
    if (t == 0)
        for (int k = n - 1; k >= 0; --k) {
            const int p = perm[k];
            const int a = s_idx[k]; s_idx[k] = s_idx[p]; s_idx[p] = a;
        }
    __syncthreads();
    for (int j = t; j < n; j += blockDim.x) idx[j] = s_idx[j];
    for (int j = wid; j < n; j += nw) {
        const double* src = s_W + (size_t)s_idx[j] * n;
        double* dst = Out + (size_t)j * n;
        for (int i = lane; i < n; i += 32) dst[i] = src[i];
    }
}

// value of det / log_det.  policy 2 (LLT): from the Cholesky diagonal; else from the elimination stats.
// out[0] = value, out[1] = 1 if the decomposition is valid (det.hpp:58 `decomp_.valid()`)
__global__ void __launch_bounds__(1024)
det_value_kernel(int n, int policy, int take_log, const double* __restrict__ diag, const double* __restrict__ stats, double* out)
{
    __shared__ double sm[32 * 2];
    if (policy == 2) {
        double acc[2] = {0.0, 0.0};                      // sum log L_ii ; (product handled below)
        double prod = 1.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) { acc[0] += log(diag[i]); prod *= diag[i]; }
        for (int o = 16; o > 0; o >>= 1) prod *= __shfl_down_sync(0xffffffffu, prod, o);
        __shared__ double sp[32];
        if ((threadIdx.x & 31) == 0) sp[threadIdx.x >> 5] = prod;
        block_sum<2>(acc, sm);
        if (threadIdx.x == 0) {
            double pr = 1.0;
            for (int w2 = 0; w2 < (int)(blockDim.x >> 5); ++w2) pr *= sp[w2];
            const bool ok = stats[0] == 0.0;
            out[0] = !ok ? NAN : (take_log ? 2.0 * acc[0] : pr * pr);
            out[1] = ok ? 1.0 : 0.0;
        }
        return;
    }
    if (threadIdx.x == 0) {
        const double det = stats[1], la = stats[2];
        const double v = take_log ? la : det;
        bool ok = stats[0] == 0.0;
        if (policy == 0) ok = ok && stats[3] > 2.220446049250313e-16 * (double)n * stats[4];   // FullPivLU::isInvertible
        else ok = ok && (take_log ? isfinite(la) : det != 0.0);                               // det.hpp:143, log_det.hpp:143
        out[0] = (stats[0] != 0.0 && !take_log) ? 0.0 : (stats[0] != 0.0 ? -INFINITY : v);
        out[1] = ok ? 1.0 : 0.0;
    }
}

// A_adj += f * inv^T, f = seed (log_det) or seed * det (det.hpp:61-62, log_det.hpp:61-62)
__global__ void det_adj_kernel(int n, const double* __restrict__ inv, const double* __restrict__ val, int take_log, double seed,
                               int overwrite, double* A_adj)
{
    if (val[1] == 0.0) return;
    const size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (e >= (size_t)n * n) return;
    const int i = (int)(e % n), j = (int)(e / n);
    const double f = take_log ? seed : seed * val[0];
    const double g = f * inv[j + (size_t)i * n];
    A_adj[e] = overwrite ? g : A_adj[e] + g;
}

// wishart: s[0] = log det L_v, s[1] = log det L_x, s[2] = invalid flag
__global__ void __launch_bounds__(1024)
ws_logdet_kernel(int n, const double* __restrict__ diag, const double* __restrict__ flag, double* s, int slot)
{
    __shared__ double sm[32];
    double acc[1] = {0.0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc[0] += log(diag[i]);
    block_sum<1>(acc, sm);
    if (threadIdx.x == 0) {
        s[slot] = acc[0];
        if (slot == 0) s[2] = 0.0;
        if (flag[0] != 0.0) s[2] = 1.0;
    }
}
__global__ void __launch_bounds__(1024)
ws_value_kernel(int p, double df, const double* __restrict__ G1, double* s, double* out)
{
    __shared__ double sm[32];
    double acc[1] = {0.0};
    for (int i = threadIdx.x; i < p; i += blockDim.x) acc[0] += G1[i + (size_t)i * p];
    block_sum<1>(acc, sm);
    if (threadIdx.x == 0) {
        const bool ok = s[2] == 0.0 && (df + 1.0 > (double)p);                      // wishart.hpp:203-207
        if (!ok) s[2] = 1.0;
        out[0] = ok ? (df - (double)p - 1.0) * s[1] - 0.5 * acc[0] - df * s[0] : -INFINITY;   // wishart.hpp:157-165
    }
}
__global__ void ws_adj_kernel(int p, double df, double seed, int overwrite, const double* __restrict__ s,
                              const double* __restrict__ Xinv, const double* __restrict__ Vinv, const double* __restrict__ G2,
                              double* X_adj, double* V_adj)
{
    if (s[2] != 0.0) return;                                                        // wishart.hpp:170
    const size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (e >= (size_t)p * p) return;
    const double pd = (double)p;
    if (V_adj) {
        const double g = (0.5 * seed) * (G2[e] - df * Vinv[e]);
        V_adj[e] = overwrite ? g : V_adj[e] + g;
    }
    if (X_adj) {
        const double g = (0.5 * seed) * ((df - pd - 1.0) * Xinv[e] - Vinv[e]);
        X_adj[e] = overwrite ? g : X_adj[e] + g;
    }
}

struct WishartWs { double *Vinv = nullptr, *Xinv = nullptr, *G1 = nullptr, *G2 = nullptr, *s = nullptr; size_t cap = 0; };
static WishartWs g_wws[kMaxDevices];
struct GjWs { int *perm = nullptr, *idx = nullptr; double* val = nullptr; size_t cap = 0; };
static GjWs g_gj[kMaxDevices];

static ShutdownHook g_dense_release([](int d) {
    DenseWs& w = g_dws[d];
    for (double* p : {w.L, w.Y, w.X, w.d, w.z, w.scal, w.diag}) cudaFree(p);
    w = DenseWs();
    cudaFree(g_trplan[d].dev);
    g_trplan[d] = TrPlan();
    WishartWs& ws = g_wws[d];
    for (double* p : {ws.Vinv, ws.Xinv, ws.G1, ws.G2, ws.s}) cudaFree(p);
    ws = WishartWs();
    GjWs& g = g_gj[d];
    cudaFree(g.perm); cudaFree(g.idx); cudaFree(g.val);
    g = GjWs();
});

static int ensure_gj(Context& c, size_t n, GjWs** out)
{
    GjWs& g = g_gj[c.device];
    if (g.cap < n) {
        FADB_CUDA(cudaStreamSynchronize(c.stream));
        if (g.perm) FADB_CUDA(cudaFree(g.perm));
        if (g.idx) FADB_CUDA(cudaFree(g.idx));
        FADB_CUDA(cudaMalloc(&g.perm, n * sizeof(int)));
        FADB_CUDA(cudaMalloc(&g.idx, n * sizeof(int)));
        if (!g.val) FADB_CUDA(cudaMalloc(&g.val, 2 * sizeof(double)));
        FADB_CUDA(cudaFuncSetAttribute(gj_colmap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        g.cap = n;
    }
    *out = &g;
    return FADB_OK;
}
static int ensure_wws(Context& c, size_t p, WishartWs** out)
{
    WishartWs& w = g_wws[c.device];
    if (w.cap < p) {
        FADB_CUDA(cudaStreamSynchronize(c.stream));
        for (double* q : {w.Vinv, w.Xinv, w.G1, w.G2}) if (q) FADB_CUDA(cudaFree(q));
        for (double** q : {&w.Vinv, &w.Xinv, &w.G1, &w.G2}) FADB_CUDA(cudaMalloc(q, p * p * sizeof(double)));
        if (!w.s) FADB_CUDA(cudaMalloc(&w.s, 4 * sizeof(double)));
        w.cap = p;
    }
    *out = &w;
    return FADB_OK;
}

// general inverse + pivot statistics by Gauss-Jordan: w->X = A^-1, w->scal = stats
static int gj_inverse(Context* c, DenseWs* w, GjWs* g, int N, const double* A)
{
    const size_t nn = (size_t)N * N;
    cudaStream_t st = c->stream;
    FADB_CUDA(cudaMemsetAsync(w->scal, 0, 8 * sizeof(double), st));
    static const bool small_ok = [] { const char* e = getenv("FADB_GJ_SMALL"); return !(e && e[0] == '0'); }();   // =0: A/B runs
    if (small_ok && N <= GJ_SMALL_MAX) {
        static bool attr_set[kMaxDevices] = {false};
        if (!attr_set[c->device]) {
            FADB_CUDA(cudaFuncSetAttribute(gj_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(GJ_SMALL_MAX * GJ_SMALL_MAX * sizeof(double))));
            attr_set[c->device] = true;
        }
        gj_small_kernel<<<1, 1024, nn * sizeof(double), st>>>(N, A, w->d, w->z, g->perm, w->scal, g->idx, w->X);
        c->launches++;
        FADB_CUDA(cudaGetLastError());
        return FADB_OK;
    }
    FADB_CUDA(cudaMemcpyAsync(w->L, A, nn * sizeof(double), cudaMemcpyDeviceToDevice, st));
    dim3 ugrid((N + 127) / 128, (N + 15) / 16);
    for (int k = 0; k < N; ++k) {
        gj_pivot_kernel<<<1, 1024, 0, st>>>(N, w->L, k, w->d, w->z, g->perm, w->scal);
        gj_update_kernel<<<ugrid, 128, 0, st>>>(N, w->L, k, w->d, w->z);
        c->launches += 2;
    }
    gj_colmap_kernel<<<1, 1024, (size_t)N * sizeof(int), st>>>(N, g->perm, g->idx);
    gj_gather_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, st>>>(N, w->L, g->idx, w->X);
    c->launches += 2;
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
        rc = spd_factor_inverse(c, w, tp, N, Sigma);
        if (rc) return rc;
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

// det / log_det of a square matrix and its adjoint A_adj += seed * [det] * A^{-T}.  policy:
// FADB_DET_FULLPIVLU | FADB_DET_LDLT | FADB_DET_LLT.  factor = 0 reuses the decomposition of the
// previous call on this device (backward pass after a forward-only evaluation).
extern "C" int fadb_det_async(size_t n, const double* A, double* A_adj, double seed, unsigned flags, int policy, int take_log,
                              int factor, double* dev_value)
{
    FADB_REQUIRE(n >= 1 && n <= 46340, "fadb_det: n out of range");
    FADB_REQUIRE(A && dev_value, "fadb_det: null pointer");
    FADB_REQUIRE(policy >= 0 && policy <= 2, "fadb_det: unknown decomposition policy");
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    DenseWs* w;
    if ((rc = ensure_dws(*c, n, &w))) return rc;
    GjWs* g;
    if ((rc = ensure_gj(*c, n, &g))) return rc;
    const int N = (int)n;
    cudaStream_t st = c->stream;
    if (factor) {
        if (policy == 2) {
            TrPlan* tp;
            if ((rc = ensure_trplan(*c, n, &tp))) return rc;
            if ((rc = spd_factor_inverse(c, w, tp, N, A))) return rc;
        } else {
            if ((rc = gj_inverse(c, w, g, N, A))) return rc;
        }
        det_value_kernel<<<1, 1024, 0, st>>>(N, policy, take_log, w->diag, w->scal, g->val);
        c->launches++;
    }
    FADB_CUDA(cudaMemcpyAsync(dev_value, g->val, sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (A_adj && seed != 0.0) {
        const size_t nn = n * n;
        det_adj_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, st>>>(N, w->X, g->val, take_log, seed,
                                                                      (flags & FADB_OVERWRITE_ADJ) ? 1 : 0, A_adj);
        c->launches++;
    }
    FADB_CUDA(cudaGetLastError());
    return FADB_OK;
}

extern "C" int fadb_det(size_t n, const double* A, double* A_adj, double seed, unsigned flags, int policy, int take_log,
                        double* host_value)
{
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    double* dval = c->scalars + 33;
    rc = fadb_det_async(n, A, A_adj, seed, flags, policy, take_log, 1, dval);
    if (rc) return rc;
    if (host_value) {
        FADB_CUDA(cudaMemcpyAsync(c->host_scalars + 6, dval, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        FADB_CUDA(cudaStreamSynchronize(c->stream));
        *host_value = c->host_scalars[6];
    }
    return FADB_OK;
}

// wishart_adj_log_pdf(X, V, df) (reverse/stat/wishart.hpp:140-214): Cholesky + inverse of V and X (lower
// triangles), value (df-p-1) log det L_x - 0.5 tr(X V^-1) - df log det L_v, adjoints
// X_adj += 0.5 seed ((df-p-1) X^-1 - V^-1), V_adj += 0.5 seed (V^-1 X V^-1 - df V^-1); -inf and no adjoints
// when either matrix is not positive definite or df + 1 <= p.  X V^-1 uses the FULL matrix X, as the reference.
extern "C" int fadb_wishart_async(size_t p, const double* X, const double* V, double df, double* X_adj, double* V_adj,
                                  double seed, unsigned flags, int factor, double* dev_value)
{
    FADB_REQUIRE(p >= 1 && p <= 46340, "fadb_wishart: p out of range");
    FADB_REQUIRE(X && V && dev_value, "fadb_wishart: null pointer");
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    DenseWs* w;
    if ((rc = ensure_dws(*c, p, &w))) return rc;
    WishartWs* ww;
    if ((rc = ensure_wws(*c, p, &ww))) return rc;
    TrPlan* tp;
    if ((rc = ensure_trplan(*c, p, &tp))) return rc;
    const int P = (int)p;
    const size_t pp = p * p;
    cudaStream_t st = c->stream;
    if (factor) {
        if ((rc = spd_factor_inverse(c, w, tp, P, V))) return rc;
        ws_logdet_kernel<<<1, 1024, 0, st>>>(P, w->diag, w->scal, ww->s, 0);
        FADB_CUDA(cudaMemcpyAsync(ww->Vinv, w->X, pp * sizeof(double), cudaMemcpyDeviceToDevice, st));
        if ((rc = spd_factor_inverse(c, w, tp, P, X))) return rc;
        ws_logdet_kernel<<<1, 1024, 0, st>>>(P, w->diag, w->scal, ww->s, 1);
        FADB_CUDA(cudaMemcpyAsync(ww->Xinv, w->X, pp * sizeof(double), cudaMemcpyDeviceToDevice, st));
        gemm<false, false>(*c, P, P, P, 1.0, X, P, ww->Vinv, P, 0.0, ww->G1, P, 0, 0);          // xv_inv_, wishart.hpp:197
        ws_value_kernel<<<1, 1024, 0, st>>>(P, df, ww->G1, ww->s, ww->s + 3);
        c->launches += 3;
    }
    FADB_CUDA(cudaMemcpyAsync(dev_value, ww->s + 3, sizeof(double), cudaMemcpyDeviceToDevice, st));
    if ((X_adj || V_adj) && seed != 0.0) {
        if (V_adj) gemm<false, false>(*c, P, P, P, 1.0, ww->Vinv, P, ww->G1, P, 0.0, ww->G2, P, 0, 0);
        ws_adj_kernel<<<(unsigned)((pp + 255) / 256), 256, 0, st>>>(P, df, seed, (flags & FADB_OVERWRITE_ADJ) ? 1 : 0, ww->s,
                                                                     ww->Xinv, ww->Vinv, ww->G2, X_adj, V_adj);
        c->launches++;
    }
    FADB_CUDA(cudaGetLastError());
    return FADB_OK;
}

extern "C" int fadb_wishart(size_t p, const double* X, const double* V, double df, double* X_adj, double* V_adj, double seed,
                            unsigned flags, double* host_value)
{
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    double* dval = c->scalars + 34;
    rc = fadb_wishart_async(p, X, V, df, X_adj, V_adj, seed, flags, 1, dval);
    if (rc) return rc;
    if (host_value) {
        FADB_CUDA(cudaMemcpyAsync(c->host_scalars + 7, dval, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        FADB_CUDA(cudaStreamSynchronize(c->stream));
        *host_value = c->host_scalars[7];
    }
    return FADB_OK;
}
