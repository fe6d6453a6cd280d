// p2p.cu — the small all-reduce of the sharded reductions over NVLink peer memory.
//
// Element-range sharding leaves ONE exchange per evaluation: a handful of partial sums
// (1 double for sum(f(x)), 4 for normal_adj_log_pdf, cols + 1 for the regression, 26 for the network).  At
// that size a collective is pure latency: NCCL's all-reduce costs 15-25 us per call on this box against
// 60-800 us of kernel time.  Here every rank owns a mailbox in its HBM, exported to the other ranks of the
// node with CUDA IPC; the all-reduce is one tiny kernel per rank:
//   push   my n doubles into entry [my rank] of EVERY rank's mailbox (peer stores through NVLink / NVSwitch); each double
//          travels as two 8-byte words {32 data bits | 32 sequence bits}: the flag is IN the payload (p2p_device.cuh),
//          no fence, no separate flag store;
//   wait   until the words of MY mailbox carry this sequence number (bounded spin per element);
//   reduce the world's entries in rank order (every rank forms the bit-identical sum) into the caller's buffer.
// Two payload buffers alternate by sequence parity: a rank can only start evaluation k + 2 after all ranks
// raised flag k + 1, i.e. after they finished reading buffer k.  No NCCL, no host synchronisation.
#include <cstring>

#include "common.cuh"
#include "p2p_device.cuh"

namespace fadb {

struct P2PState {
    int world = 0, rank = -1;
    bool ipc = false;                        // peers opened through CUDA IPC (one process per GPU) or plain peer access (one process)
    Mailbox* local = nullptr;
    Mailbox* peer[P2P_MAXW] = {};
    unsigned long long seq = 0;
    unsigned long long timeout_ns = 0;
    unsigned long long* host_error = nullptr;        // page-locked + mapped: written by the device on a timeout
    unsigned long long* host_error_dev = nullptr;    // its device alias
    cudaIpcMemHandle_t handle;
};
static P2PState g_p2p[kMaxDevices];

static unsigned long long timeout_from_env()
{
    const char* e = getenv("FADB_P2P_TIMEOUT_MS");
    const double ms = e ? atof(e) : 30000.0;         // generous: a peer may sit in a first-call NVRTC compile or a module load
    return (unsigned long long)((ms > 0 ? ms : 30000.0) * 1e6);
}

// release everything of device d's exchange state (current device already set)
static void p2p_teardown(P2PState& s)
{
    if (!s.local) return;
    if (s.ipc)
        for (int r = 0; r < s.world; ++r)
            if (r != s.rank && s.peer[r]) cudaIpcCloseMemHandle(s.peer[r]);
    cudaFree(s.local);
    if (s.host_error) cudaFreeHost(s.host_error);
    s = P2PState();
}
static ShutdownHook g_p2p_release([](int d) { p2p_teardown(g_p2p[d]); });

static int p2p_alloc(P2PState& s, int world, int rank)
{
    FADB_CUDA(cudaMalloc(&s.local, sizeof(Mailbox)));
    FADB_CUDA(cudaMemset(s.local, 0, sizeof(Mailbox)));
    FADB_CUDA(cudaHostAlloc(&s.host_error, sizeof(unsigned long long), cudaHostAllocMapped | cudaHostAllocPortable));
    *s.host_error = 0;
    FADB_CUDA(cudaHostGetDevicePointer(&s.host_error_dev, s.host_error, 0));
    FADB_CUDA(cudaDeviceSynchronize());
    s.world = world; s.rank = rank; s.seq = 0; s.timeout_ns = timeout_from_env();
    return FADB_OK;
}

// A timed-out exchange is sticky: every later host entry of the exchange fails instead of returning poisoned numbers.
static int p2p_check_healthy(const P2PState& s)
{
    if (s.host_error && *reinterpret_cast<volatile unsigned long long*>(s.host_error)) {
        set_error("fastad_b200: peer-memory exchange " + std::to_string(*s.host_error) +
                  " timed out (a rank did not arrive: dead peer or mismatched rank count); results since then are NaN. "
                  "fadb_p2p_destroy + fadb_p2p_create to recover");
        return FADB_ERR_STATE;
    }
    return FADB_OK;
}

struct P2PArgs { P2PLink link; int n; double* buf; };

__global__ void __launch_bounds__(P2P_MAXN)
p2p_allreduce_kernel(P2PArgs a)
{
    const P2PLink& l = a.link;
    const int t = threadIdx.x;
    // peer stores (NVLink; r == rank is local): consecutive threads write consecutive 16-byte entries of one mailbox
    for (int idx = t; idx < a.n * l.world; idx += blockDim.x) {
        const int r = idx / a.n, k = idx - r * a.n;
        p2p_send(l, r, k, a.buf[k]);
    }
    __syncthreads();                 // every thread has read buf[] before any thread overwrites it
    if (t < a.n) a.buf[t] = p2p_pull(l, t, a.n);
}

static void fill_link(const P2PState& s, P2PLink* out)
{
    for (int r = 0; r < P2P_MAXW; ++r) out->peer[r] = s.peer[r];
    out->local = s.local; out->world = s.world; out->rank = s.rank; out->seq = s.seq;
    out->timeout_ns = s.timeout_ns; out->host_error = s.host_error_dev;
}
int p2p_next_link(Context& c, P2PLink* out)
{
    P2PState& s = g_p2p[c.device];
    if (!s.local || !s.peer[s.world - 1]) { set_error("fastad_b200: peer-memory exchange is not connected (fadb_p2p_create / fadb_p2p_connect)"); return FADB_ERR_STATE; }
    int rc = p2p_check_healthy(s);
    if (rc) return rc;
    ++s.seq;
    fill_link(s, out);
    return FADB_OK;
}
int p2p_current_link(Context& c, P2PLink* out)
{
    P2PState& s = g_p2p[c.device];
    if (!s.local || !s.peer[s.world - 1] || s.seq == 0) { set_error("fastad_b200: no pushed exchange to pull"); return FADB_ERR_STATE; }
    int rc = p2p_check_healthy(s);
    if (rc) return rc;
    fill_link(s, out);
    return FADB_OK;
}

}  // namespace fadb

using namespace fadb;

// Allocate this rank's mailbox on the current device and return its 64-byte CUDA IPC handle, to be exchanged
// with the other ranks by the caller (torch.distributed.all_gather, MPI_Allgather, ...).  Calling it again
// releases the previous mailbox and the peers opened from it first.
extern "C" int fadb_p2p_create(int world, int rank, void* out_handle64)
{
    FADB_REQUIRE(world >= 1 && world <= P2P_MAXW && rank >= 0 && rank < world && out_handle64, "fadb_p2p_create: bad arguments");
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    P2PState& s = g_p2p[c->device];
    if (s.local) { FADB_CUDA(cudaStreamSynchronize(c->stream)); p2p_teardown(s); }
    if ((rc = p2p_alloc(s, world, rank))) return rc;
    s.ipc = true;
    FADB_CUDA(cudaIpcGetMemHandle(&s.handle, s.local));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(out_handle64, &s.handle, 64);
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
    FADB_REQUIRE(s.local != nullptr && s.ipc, "fadb_p2p_connect: call fadb_p2p_create first");
    for (int r = 0; r < s.world; ++r) {
        if (r == s.rank) { s.peer[r] = s.local; continue; }
        if (s.peer[r]) continue;     // already open
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(all_handles) + 64 * (size_t)r, 64);
        void* p = nullptr;
        FADB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        s.peer[r] = static_cast<Mailbox*>(p);
    }
    return FADB_OK;
}

// ONE process driving several GPUs: rank r = devices[r]; mailboxes are allocated on every
// listed device and connected through plain peer access (cudaDeviceEnablePeerAccess) -- no IPC, no handle exchange.
// Afterwards the caller selects a device (cudaSetDevice / fadb_init) and issues that rank's calls as usual; a rank's
// pull must be issued after every rank's push has been ENQUEUED (the host thread enqueues them one after the other).
extern "C" int fadb_p2p_create_local(int world, const int* devices)
{
    FADB_REQUIRE(world >= 1 && world <= P2P_MAXW && devices, "fadb_p2p_create_local: bad arguments");
    int prev = 0;
    FADB_CUDA(cudaGetDevice(&prev));
    for (int r = 0; r < world; ++r) {
        FADB_REQUIRE(devices[r] >= 0 && devices[r] < kMaxDevices, "fadb_p2p_create_local: bad device index");
        for (int q = 0; q < r; ++q) FADB_REQUIRE(devices[q] != devices[r], "fadb_p2p_create_local: a device is listed twice");
    }
    for (int r = 0; r < world; ++r) {
        FADB_CUDA(cudaSetDevice(devices[r]));
        Context* c;
        int rc = current_context(&c);
        if (rc) return rc;
        for (int q = 0; q < world; ++q) {
            if (q == r) continue;
            int can = 0;
            FADB_CUDA(cudaDeviceCanAccessPeer(&can, devices[r], devices[q]));
            FADB_REQUIRE(can, "fadb_p2p_create_local: the devices have no peer access to each other");
            cudaError_t e = cudaDeviceEnablePeerAccess(devices[q], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else FADB_CUDA(e);
        }
        P2PState& s = g_p2p[devices[r]];
        if (s.local) { FADB_CUDA(cudaStreamSynchronize(c->stream)); p2p_teardown(s); }
        if ((rc = p2p_alloc(s, world, r))) return rc;
        s.ipc = false;
    }
    for (int r = 0; r < world; ++r)
        for (int q = 0; q < world; ++q) g_p2p[devices[r]].peer[q] = g_p2p[devices[q]].local;
    FADB_CUDA(cudaSetDevice(prev));
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
    P2PArgs a;
    if ((rc = p2p_next_link(*c, &a.link))) return rc;
    a.n = n; a.buf = dev_buf;
    p2p_allreduce_kernel<<<1, P2P_MAXN, 0, c->stream>>>(a);
    c->launches++;
    FADB_CUDA(cudaGetLastError());
    return FADB_OK;
}

// 0 = healthy; otherwise the sequence number at which a wait timed out (a peer did not arrive).  Synchronises the stream.
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

// Release the current device's mailbox.  Peers store into it during THEIR pushes, so every rank must have finished its
// last exchange first: the caller puts a barrier (torch.distributed.barrier, MPI_Barrier) between the last evaluation
// and this call.
extern "C" int fadb_p2p_destroy(void)
{
    Context* c;
    int rc = current_context(&c);
    if (rc) return rc;
    P2PState& s = g_p2p[c->device];
    if (!s.local) return FADB_OK;
    FADB_CUDA(cudaStreamSynchronize(c->stream));
    p2p_teardown(s);
    return FADB_OK;
}
