// common.cuh — shared device/host helpers for the fastad_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/fastad_b200.h"

namespace fadb {

// ---------------------------------------------------------------------------------------
// Error convention: every C-ABI entry returns 0 or a negative code; message is thread-local.
// No C++ exception crosses the ABI.
// ---------------------------------------------------------------------------------------
void set_error(const std::string& msg);
// Modules that keep lazily allocated per-device workspaces register a release function; fadb_shutdown calls each
// with every initialised device (current device already set, library stream already synchronised).
void on_shutdown(void (*release)(int device));
struct ShutdownHook { explicit ShutdownHook(void (*release)(int device)) { on_shutdown(release); } };
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define FADB_CUDA(expr)                                                             \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) return ::fadb::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define FADB_REQUIRE(cond, msg)                                  \
    do {                                                         \
        if (!(cond)) {                                           \
            ::fadb::set_error(std::string("fastad_b200: ") + msg); \
            return FADB_ERR_ARG;                                 \
        }                                                        \
    } while (0)

// ---------------------------------------------------------------------------------------
// Per-device context: SM count, the stream every launch goes to, and a small workspace
// (per-CTA partials + ticket counters) so that hot calls never allocate.
// ---------------------------------------------------------------------------------------
constexpr int kMaxGrid = 148 * 16;       // upper bound on CTAs of any reduction launch
constexpr int kPartialSlots = 8;         // reduction channels per CTA
constexpr int kMaxDevices = 16;
constexpr int kScalarSlots = 256;        // device doubles for reduced values / deferred seeds

struct Context {
    bool ready = false;
    int device = 0;
    int sms = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    double* partials = nullptr;          // [kMaxGrid * kPartialSlots]
    unsigned int* tickets = nullptr;     // [64] last-block tickets (self-resetting)
    double* scalars = nullptr;           // [kScalarSlots] device result slots
    double* host_scalars = nullptr;      // pinned mirror for D2H of results
    unsigned long long launches = 0;     // kernels launched by this library on this device
};

int current_context(Context** out);

inline int grid_for(const Context& c, size_t work_items, int per_sm, size_t items_per_cta)
{
    size_t want = (work_items + items_per_cta - 1) / items_per_cta;
    size_t cap = (size_t)c.sms * per_sm;
    if (want < 1) want = 1;
    if (want > cap) want = cap;
    if (want > (size_t)kMaxGrid) want = kMaxGrid;
    return (int)want;
}

// One resident wave: grid = SMs x (CTAs that fit per SM for this kernel), capped by the work.
// A grid-stride kernel launched like this has no second-wave tail.
template <class Kernel>
inline int resident_grid(const Context& c, Kernel kernel, int threads, size_t dyn_smem,
                         size_t work_items, size_t items_per_cta)
{
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, dyn_smem) != cudaSuccess ||
        per_sm < 1)
        per_sm = 1;
    return grid_for(c, work_items, per_sm, items_per_cta);
}

// 256-bit global memory access helpers (double4_t, ldg256_stream, ld256, st256): device_reduce.cuh
__host__ __device__ inline bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }

}  // namespace fadb
#include "device_reduce.cuh"   // warp_sum, warp_prod, block_sum
namespace fadb {

// Grid-level finish: every CTA stores its NCH partials; the last CTA to arrive (ticket)
// sums all partials in CTA order with one warp-strided pass + fixed tree, then calls `fin`
// from thread 0.  One launch, deterministic, no float atomics.
template <int NCH, class Fin>
__device__ __forceinline__ void grid_finish(double (&acc)[NCH], double* smem, double* partials,
                                            unsigned int* ticket, Fin&& fin)
{
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) partials[(size_t)blockIdx.x * NCH + c] = acc[c];
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
        if (is_last) __threadfence();      // acquire side, once: the barrier below orders the other threads' ld.cg reads after it
    }
    __syncthreads();
    if (!is_last) return;
    double tot[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) tot[c] = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) tot[c] += __ldcg(&partials[(size_t)b * NCH + c]);
    }
    block_sum<NCH>(tot, smem);
    if (threadIdx.x == 0) {
        *ticket = 0u;   // self-reset for the next launch on this stream
        fin(tot);
    }
}

// ---- programmatic dependent launch (PDL).  Back-to-back launches of the full-GPU, single-wave
// kernels otherwise pay the launch latency and the idle tail of the previous grid every time.
// launch_pdl() marks the launch as allowed to become resident while the previous kernel on the
// stream is still running; the kernel calls pdl_prologue() before its FIRST global-memory access,
// which (a) lets the launch after it do the same and (b) blocks until the previous grid has
// completed and flushed.  Ordering of memory effects is therefore unchanged.  FADB_NO_PDL=1 disables.
__device__ __forceinline__ void pdl_prologue()
{
    cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();
}
inline bool pdl_enabled()
{
    static const bool on = [] { const char* e = getenv("FADB_NO_PDL"); return !(e && e[0] == '1'); }();   // thread-safe on first use
    return on;
}
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace fadb
