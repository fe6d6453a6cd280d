// p2p.cu — the small all-reduce of the sharded reductions over NVLink peer memory.
//
// Element-range sharding leaves ONE exchange per evaluation: a handful of partial sums
// (1 double for sum(f(x)), 4 for normal_adj_log_pdf, cols + 1 for the regression, 26 for the network).  At
// that size a collective is pure latency: NCCL's all-reduce costs 15-25 us per call on this box against
// 60-800 us of kernel time.  Here every rank owns a mailbox in its HBM, exported to the other ranks of the
// node with CUDA IPC; the all-reduce is one tiny kernel per rank:
//   push   my n doubles into slot[my rank] of EVERY rank's mailbox (peer stores through NVLink / NVSwitch),
//          fence (system scope), then raise flag[my rank] = sequence number in every mailbox;
//   wait   until all flags of MY mailbox carry this sequence number (bounded spin);
//   reduce the world's slots in rank order (every rank forms the bit-identical sum) into the caller's buffer.
// Two payload buffers alternate by sequence parity: a rank can only start evaluation k + 2 after all ranks
// raised flag k + 1, i.e. after they finished reading buffer k.  No NCCL, no host synchronisation.
#include <cstring>

#include "common.cuh"
#include "p2p_device.cuh"

namespace fadb {

struct P2PState {
    int world = 0, rank = -1;
    Mailbox* local = nullptr;
    Mailbox* peer[P2P_MAXW] = {};
    unsigned long long seq = 0;
    cudaIpcMemHandle_t handle;
};
static P2PState g_p2p[kMaxDevices];
static ShutdownHook g_p2p_release([](int d) {
    P2PState& s = g_p2p[d];
    if (!s.local) return;
    for (int r = 0; r < s.world; ++r)
        if (r != s.rank && s.peer[r]) cudaIpcCloseMemHandle(s.peer[r]);
    cudaFree(s.local);
    s = P2PState();
});

struct P2PArgs { Mailbox* peer[P2P_MAXW]; Mailbox* local; int world, rank, n; unsigned long long seq; double* buf; };

__global__ void __launch_bounds__(P2P_MAXN)
p2p_allreduce_kernel(P2PArgs a)
{
    const int t = threadIdx.x, b = (int)(a.seq & 1);
    // peer stores (NVLink; r == rank is local): consecutive threads write consecutive doubles of one mailbox
    for (int idx = t; idx < a.n * a.world; idx += blockDim.x) {
        const int r = idx / a.n, k = idx - r * a.n;
        a.peer[r]->slot[b][a.rank][k] = a.buf[k];
    }
    __threadfence_system();
    __syncthreads();
    if (t < a.world) {
        __threadfence_system();      // release by the flag's writer (cumulative over the stores ordered before the barrier)
        *reinterpret_cast<volatile unsigned long long*>(&a.peer[t]->flag[a.rank]) = a.seq;
        volatile unsigned long long* f = &a.local->flag[t];
        const long long t0 = clock64();
        while (*f < a.seq) {
            if (clock64() - t0 > (1ll << 33)) { a.local->error = a.seq; break; }      // ~4 s at 2 GHz: a peer died
        }
    }
    __threadfence_system();
    __syncthreads();
    if (t < a.n) {
        double s = 0.0;
        for (int r = 0; r < a.world; ++r) s += *reinterpret_cast<volatile double*>(&a.local->slot[b][r][t]);
        a.buf[t] = s;
    }
}

static void fill_link(const P2PState& s, P2PLink* out)
{
    for (int r = 0; r < P2P_MAXW; ++r) out->peer[r] = s.peer[r];
    out->local = s.local; out->world = s.world; out->rank = s.rank; out->seq = s.seq;
}
int p2p_next_link(Context& c, P2PLink* out)
{
    P2PState& s = g_p2p[c.device];
    if (!s.local || !s.peer[s.world - 1]) { set_error("fastad_b200: peer-memory exchange is not connected (fadb_p2p_create / fadb_p2p_connect)"); return FADB_ERR_STATE; }
    ++s.seq;
    fill_link(s, out);
    return FADB_OK;
}
int p2p_current_link(Context& c, P2PLink* out)
{
    P2PState& s = g_p2p[c.device];
    if (!s.local || !s.peer[s.world - 1] || s.seq == 0) { set_error("fastad_b200: no pushed exchange to pull"); return FADB_ERR_STATE; }
    fill_link(s, out);
    return FADB_OK;
}

}  // namespace fadb

using namespace fadb;

// Allocate this rank's mailbox on the current device and return its 64-byte CUDA IPC handle, to be exchanged
// with the other ranks by the caller (torch.distributed.all_gather, MPI_Allgather, ...).
extern "C" int fadb_p2p_create(int world, int rank, void* out_handle64)
{
    FADB_REQUIRE(world >= 1 && world <= P2P_MAXW && rank >= 0 && rank < world && out_handle64, "fadb_p2p_create: bad arguments");
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    P2PState& s = g_p2p[c->device];
    if (s.local) { FADB_CUDA(cudaFree(s.local)); s = P2PState(); }
    FADB_CUDA(cudaMalloc(&s.local, sizeof(Mailbox)));
    FADB_CUDA(cudaMemset(s.local, 0, sizeof(Mailbox)));
    FADB_CUDA(cudaDeviceSynchronize());
    FADB_CUDA(cudaIpcGetMemHandle(&s.handle, s.local));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(out_handle64, &s.handle, 64);
    s.world = world; s.rank = rank;
    return FADB_OK;
}

// all_handles: world x 64 bytes in rank order (this rank's own entry is ignored).
extern "C" int fadb_p2p_connect(const void* all_handles)
{
    FADB_REQUIRE(all_handles != nullptr, "fadb_p2p_connect: null handles");
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    P2PState& s = g_p2p[c->device];
    FADB_REQUIRE(s.local != nullptr, "fadb_p2p_connect: call fadb_p2p_create first");
    for (int r = 0; r < s.world; ++r) {
        if (r == s.rank) { s.peer[r] = s.local; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(all_handles) + 64 * (size_t)r, 64);
        void* p = nullptr;
        FADB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        s.peer[r] = static_cast<Mailbox*>(p);
    }
    return FADB_OK;
}

// In-place sum of dev_buf[0..n) over the ranks, asynchronous on the library stream.  Every rank must call it
// the same number of times with the same n.
extern "C" int fadb_p2p_allreduce(double* dev_buf, int n)
{
    FADB_REQUIRE(dev_buf && n >= 1 && n <= P2P_MAXN, "fadb_p2p_allreduce: n must be in 1..128");
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    P2PState& s = g_p2p[c->device];
    FADB_REQUIRE(s.local && s.peer[s.world - 1], "fadb_p2p_allreduce: not connected");
    P2PArgs a;
    for (int r = 0; r < P2P_MAXW; ++r) a.peer[r] = s.peer[r];
    a.local = s.local; a.world = s.world; a.rank = s.rank; a.n = n; a.seq = ++s.seq; a.buf = dev_buf;
    p2p_allreduce_kernel<<<1, P2P_MAXN, 0, c->stream>>>(a);
    c->launches++;
    FADB_CUDA(cudaGetLastError());
    return FADB_OK;
}

// 0 = healthy; otherwise the sequence number at which a wait timed out (a peer did not arrive).
extern "C" int fadb_p2p_status(unsigned long long* out_error_seq)
{
    FADB_REQUIRE(out_error_seq != nullptr, "fadb_p2p_status: null pointer");
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    P2PState& s = g_p2p[c->device];
    FADB_REQUIRE(s.local != nullptr, "fadb_p2p_status: no mailbox");
    FADB_CUDA(cudaStreamSynchronize(c->stream));
    FADB_CUDA(cudaMemcpy(out_error_seq, &s.local->error, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return FADB_OK;
}

extern "C" int fadb_p2p_destroy(void)
{
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    P2PState& s = g_p2p[c->device];
    if (!s.local) return FADB_OK;
    FADB_CUDA(cudaStreamSynchronize(c->stream));
    for (int r = 0; r < s.world; ++r)
        if (r != s.rank && s.peer[r]) cudaIpcCloseMemHandle(s.peer[r]);
    cudaFree(s.local);
    s = P2PState();
    return FADB_OK;
}
