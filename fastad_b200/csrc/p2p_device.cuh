// p2p_device.cuh — device side of the peer-memory exchange (csrc/p2p.cu), for kernels that push their
// partial sums to the other ranks from their own epilogue and pull + reduce in the consumer's prologue.
#pragma once

namespace fadb {

constexpr int P2P_MAXN = 128;        // doubles per exchange
constexpr int P2P_MAXW = 16;         // ranks per node

struct Mailbox {                     // lives in device memory, one per rank
    double slot[2][P2P_MAXW][P2P_MAXN];
    unsigned long long flag[P2P_MAXW];
    unsigned long long error;        // set by a rank whose wait timed out
};

// stands in for P2PLink in the single-GPU instantiation of a kernel template (no parameter bytes, no code)
struct NoLink {};

// by-value kernel argument; world == 0: no exchange
struct P2PLink {
    Mailbox* peer[P2P_MAXW];
    Mailbox* local;
    int world, rank;
    unsigned long long seq;
};

// Called by ALL threads of the (last) CTA that holds the reduced values: thread j < n owns value v.
// Pushes v into slot[rank][j] of every mailbox, fences, raises the flags.  `bar` is the CTA barrier to use.
template <class Bar>
__device__ __forceinline__ void p2p_push(const P2PLink& l, int j, int n, double v, Bar&& bar)
{
    const int b = (int)(l.seq & 1);
    if (j < n)
        for (int r = 0; r < l.world; ++r) l.peer[r]->slot[b][l.rank][j] = v;
    __threadfence_system();
    bar();
    if (j < l.world) {
        __threadfence_system();      // release by the flag's writer: cumulative over the payload stores ordered before bar()
        *reinterpret_cast<volatile unsigned long long*>(&l.peer[j]->flag[l.rank]) = l.seq;
    }
}

// Called by all threads of a CTA (blockDim >= world): waits for every rank's flag, then thread j < n returns
// the rank-ordered sum of slot[.][j]; other threads return 0.
__device__ __forceinline__ double p2p_pull(const P2PLink& l, int j, int n)
{
    if ((int)threadIdx.x < l.world) {
        volatile unsigned long long* f = &l.local->flag[threadIdx.x];
        const long long t0 = clock64();
        while (*f < l.seq)
            if (clock64() - t0 > (1ll << 33)) { l.local->error = l.seq; break; }      // ~4 s: a peer died
    }
    __threadfence_system();
    __syncthreads();
    double s = 0.0;
    if (j < n) {
        const int b = (int)(l.seq & 1);
        for (int r = 0; r < l.world; ++r) s += *reinterpret_cast<volatile double*>(&l.local->slot[b][r][j]);
    }
    return s;
}

// Single-thread forms for kernels whose reduced values end up in ONE thread (grid_finish epilogues, <<<1,1>>>
// finish kernels): push n <= P2P_MAXN values / wait for all ranks and return the rank-ordered sums.
__device__ __forceinline__ void p2p_push1(const P2PLink& l, const double* v, int n)
{
    const int b = (int)(l.seq & 1);
    for (int r = 0; r < l.world; ++r)
        for (int c = 0; c < n; ++c) l.peer[r]->slot[b][l.rank][c] = v[c];
    __threadfence_system();
    for (int r = 0; r < l.world; ++r) *reinterpret_cast<volatile unsigned long long*>(&l.peer[r]->flag[l.rank]) = l.seq;
}
__device__ __forceinline__ void p2p_pull1(const P2PLink& l, double* out, int n)
{
    for (int r = 0; r < l.world; ++r) {
        volatile unsigned long long* f = &l.local->flag[r];
        const long long t0 = clock64();
        while (*f < l.seq)
            if (clock64() - t0 > (1ll << 33)) { l.local->error = l.seq; break; }
    }
    __threadfence_system();
    const int b = (int)(l.seq & 1);
    for (int c = 0; c < n; ++c) {
        double s = 0.0;
        for (int r = 0; r < l.world; ++r) s += *reinterpret_cast<volatile double*>(&l.local->slot[b][r][c]);
        out[c] = s;
    }
}

// host side (p2p.cu): the link of the NEXT exchange (advances the sequence) / of the exchange just pushed
struct Context;
int p2p_next_link(Context& c, P2PLink* out);
int p2p_current_link(Context& c, P2PLink* out);

}  // namespace fadb
