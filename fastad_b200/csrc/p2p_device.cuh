// p2p_device.cuh — device side of the peer-memory exchange (csrc/p2p.cu), for kernels that push their
// partial sums to the other ranks from their own epilogue and pull + reduce in the consumer's prologue.
#pragma once

namespace fadb {

constexpr int P2P_MAXN = 128;        // doubles per exchange
constexpr int P2P_MAXW = 16;         // ranks per node

// Flag-in-payload layout (the scheme of NCCL's LL protocol): a double travels as two 8-byte words, each
// {32 data bits | low 32 bits of the exchange's sequence number << 32}.  An aligned 8-byte store is single-copy atomic, so
// a word whose flag half matches carries valid data -- the receiver needs no separate flag, the sender no fence between
// payload and flag.  (The previous protocol: payload stores, fence.sys, CTA barrier, fence.sys, flag store, flag wait:
// 8-10 us per exchange, most of it the two system-scope fences draining peer stores.)
struct Mailbox {                     // lives in device memory, one per rank
    unsigned long long ll[2][P2P_MAXW][P2P_MAXN][2];   // [sequence parity][sender][element][half]
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
// NaN once any wait on this mailbox has timed out (read after the waits of the current exchange are over)
__device__ __forceinline__ double p2p_poison(const P2PLink& l, double v)
{
    return *reinterpret_cast<volatile unsigned long long*>(&l.local->error) ? __longlong_as_double(0x7ff8000000000000ll) : v;
}

// element k of this rank's contribution -> entry [rank][k] of rank r's mailbox (peer store over NVLink; r == rank is local)
__device__ __forceinline__ void p2p_send(const P2PLink& l, int r, int k, double v)
{
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v), f = (l.seq & 0xffffffffull) << 32;
    unsigned long long* dst = l.peer[r]->ll[l.seq & 1][l.rank][k];
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"((bits & 0xffffffffull) | f), "l"((bits >> 32) | f) : "memory");
}
// element k of rank r's contribution, once both of its words carry this exchange's sequence number.  A peer that never
// arrives (dead process, wrong rank count) must not yield a plausible number: on timeout the sticky error words are set --
// the device one poisons this and every later exchange (p2p_poison), the host one makes the next host entry return
// FADB_ERR_STATE.
__device__ __forceinline__ double p2p_recv(const P2PLink& l, int r, int k)
{
    const unsigned long long* src = l.local->ll[l.seq & 1][r][k];
    const unsigned long long want = l.seq & 0xffffffffull;
    unsigned long long w0, w1, t0 = 0;
    while (true) {
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src) : "memory");
        if ((w0 >> 32) == want && (w1 >> 32) == want) break;
        if (*reinterpret_cast<volatile unsigned long long*>(&l.local->error)) break;      // already failed: do not wait again
        const unsigned long long now = p2p_now_ns();
        if (!t0) t0 = now;
        else if (now - t0 > l.timeout_ns) {
            atomicCAS(&l.local->error, 0ull, l.seq);
            if (l.host_error) { *reinterpret_cast<volatile unsigned long long*>(l.host_error) = l.seq; __threadfence_system(); }
            break;
        }
    }
    return __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
}

// Called by ALL threads of the (last) CTA that holds the reduced values: thread j < n owns value v and sends it to every
// rank.  (`bar`, the CTA barrier of the fenced protocol, is no longer needed; kept in the signature for the callers.)
template <class Bar>
__device__ __forceinline__ void p2p_push(const P2PLink& l, int j, int n, double v, Bar&&)
{
    if (j < n)
        for (int r = 0; r < l.world; ++r) p2p_send(l, (l.rank + r) % l.world, j, v);      // own mailbox first, peers staggered by rank
}

// Thread j < n returns the rank-ordered sum of element j over all ranks (NaN after a timeout); other threads return 0.
__device__ __forceinline__ double p2p_pull(const P2PLink& l, int j, int n)
{
    double s = 0.0;
    if (j < n) {
        for (int r = 0; r < l.world; ++r) s += p2p_recv(l, r, j);
        s = p2p_poison(l, s);
    }
    return s;
}

// Single-thread forms for kernels whose reduced values end up in ONE thread (grid_finish epilogues, <<<1,1>>>
// finish kernels): push n <= P2P_MAXN values / wait for all ranks and return the rank-ordered sums.
__device__ __forceinline__ void p2p_push1(const P2PLink& l, const double* v, int n)
{
    for (int r = 0; r < l.world; ++r)
        for (int c = 0; c < n; ++c) p2p_send(l, (l.rank + r) % l.world, c, v[c]);
}
__device__ __forceinline__ void p2p_pull1(const P2PLink& l, double* out, int n)
{
    for (int c = 0; c < n; ++c) {
        double s = 0.0;
        for (int r = 0; r < l.world; ++r) s += p2p_recv(l, r, c);
        out[c] = s;
    }
    for (int c = 0; c < n; ++c) out[c] = p2p_poison(l, out[c]);
}

// host side (p2p.cu): the link of the NEXT exchange (advances the sequence) / of the exchange just pushed
struct Context;
int p2p_next_link(Context& c, P2PLink* out);
int p2p_current_link(Context& c, P2PLink* out);

}  // namespace fadb
