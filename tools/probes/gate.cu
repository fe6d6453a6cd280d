// gate.cu — a host-released gate for device-side timing windows (bench.py): the stream is blocked by a one-thread kernel
// that spins on a page-locked host flag, the timed region (event, K launches, event) is ENQUEUED behind it, then the host
// opens the gate.  The K launches of a window therefore run back to back from a full queue, and the window measures the
// device time of exactly K steps instead of K steps plus the CPU's enqueue latency of the first one (which is 5 % of a
// 20 x 20 us window).  The spin gives up after 2 s so a dead host cannot hang the GPU.  Not part of the product library.
#include <cuda_runtime.h>

__global__ void gate_kernel(volatile int* flag)
{
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    while (*flag == 0) {
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        if (t - t0 > 2000000000ull) break;
    }
}

static int* g_flag = nullptr;
static int* g_flag_dev = nullptr;

extern "C" int fadb_probe_gate_arm(void* stream)
{
    if (!g_flag) {
        if (cudaHostAlloc(&g_flag, sizeof(int), cudaHostAllocMapped) != cudaSuccess) return -2;
        if (cudaHostGetDevicePointer(&g_flag_dev, g_flag, 0) != cudaSuccess) return -2;
    }
    *(volatile int*)g_flag = 0;
    gate_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(g_flag_dev);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

extern "C" int fadb_probe_gate_release(void)
{
    if (!g_flag) return -1;
    *(volatile int*)g_flag = 1;
    return 0;
}
