// p2p_device.cuh — device side of the peer-memory exchange (csrc/p2p.cu), for kernels that push their
// partial sums to the other ranks from their own epilogue and pull + reduce in the consumer's prologue.
#pragma once

namespace fadb {

constexpr int P2P_MAXN = 128;        // doubles per exchange
constexpr int P2P_MAXW = 16;         // ranks per node

struct Mailbox {                     // lives in device memory, one per rank
    double slot[2][P2P_MAXW][P2P_MAXN];
    unsigned long long flag[P2P_MAXW];
    unsigned long long error;        // STICKY: sequence number of the first exchange whose wait timed out (0 = healthy)
};

// stands in for P2PLink in the single-GPU instantiation of a kernel template (no parameter bytes, no code)
struct NoLink {};

// by-value kernel argument; world == 0: no exchange
struct P2PLink {
    Mailbox* peer[P2P_MAXW];
    Mailbox* local;
    int world, rank;
    unsigned long long seq;
    unsigned long long timeout_ns;           // bound of one wait (FADB_P2P_TIMEOUT_MS, default 30 s)
    unsigned long long* host_error;          // page-locked host word the host entry points check before every exchange
};

__device__ __forceinline__ unsigned long long p2p_now_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Wait until rank r's flag carries this exchange's sequence number.  A peer that never arrives (dead process, wrong
// rank count) must not yield a plausible number: on timeout the sticky error words are set -- the device one poisons
// this and every later exchange (p2p_poison), the host one makes the next host entry return FADB_ERR_STATE.
__device__ __forceinline__ void p2p_wait_flag(const P2PLink& l, int r)
{
    volatile unsigned long long* f = &l.local->flag[r];
    if (*f >= l.seq) return;
    const unsigned long long t0 = p2p_now_ns();
    while (*f < l.seq) {
        if (p2p_now_ns() - t0 > l.timeout_ns) {
            atomicCAS(&l.local->error, 0ull, l.seq);
            if (l.host_error) { *reinterpret_cast<volatile unsigned long long*>(l.host_error) = l.seq; __threadfence_system(); }
            break;
        }
    }
}
// NaN once any wait on this mailbox has timed out (read after the waits of the current exchange are over)
__device__ __forceinline__ double p2p_poison(const P2PLink& l, double v)
{
    return *reinterpret_cast<volatile unsigned long long*>(&l.local->error) ? __longlong_as_double(0x7ff8000000000000ll) : v;
}

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
// the rank-ordered sum of slot[.][j] (NaN after a timeout); other threads return 0.
__device__ __forceinline__ double p2p_pull(const P2PLink& l, int j, int n)
{
    if ((int)threadIdx.x < l.world) p2p_wait_flag(l, (int)threadIdx.x);
    __threadfence_system();
    __syncthreads();
    double s = 0.0;
    if (j < n) {
        const int b = (int)(l.seq & 1);
        for (int r = 0; r < l.world; ++r) s += *reinterpret_cast<volatile double*>(&l.local->slot[b][r][j]);
        s = p2p_poison(l, s);
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
    for (int r = 0; r < l.world; ++r) p2p_wait_flag(l, r);
    __threadfence_system();
    const int b = (int)(l.seq & 1);
    for (int c = 0; c < n; ++c) {
        double s = 0.0;
        for (int r = 0; r < l.world; ++r) s += *reinterpret_cast<volatile double*>(&l.local->slot[b][r][c]);
        out[c] = p2p_poison(l, s);
    }
}

// host side (p2p.cu): the link of the NEXT exchange (advances the sequence) / of the exchange just pushed
struct Context;
int p2p_next_link(Context& c, P2PLink* out);
int p2p_current_link(Context& c, P2PLink* out);

}  // namespace fadb
