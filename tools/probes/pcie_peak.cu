// pcie_peak.cu — the host-link ceiling of the end-to-end number: `reps` rounds of an H2D copy of h2d_bytes and a D2H copy of
// d2h_bytes running CONCURRENTLY (two streams, page-locked host buffers, copy engines), timed with CUDA events.
// bench.py runs it on every rank at once (barrier before), so at N GPUs the figure is the AGGREGATE host ceiling the
// fadb_black_scholes_host call competes against.  Not part of the product library.
#include <cuda_runtime.h>

extern "C" int fadb_probe_pcie(size_t h2d_bytes, size_t d2h_bytes, int reps, int mode /* 0 both, 1 h2d only, 2 d2h only */,
                               double* out_ms_per_rep)
{
    static void *h_in = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr;
    static size_t cap_in = 0, cap_out = 0;
    static cudaStream_t s0 = nullptr, s1 = nullptr;
    static cudaEvent_t e0, e1, j;
    if (!s0) {
        if (cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking) != cudaSuccess) return -2;
        if (cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking) != cudaSuccess) return -2;
        cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreateWithFlags(&j, cudaEventDisableTiming);
    }
    if (cap_in < h2d_bytes) {
        if (h_in) { cudaFreeHost(h_in); cudaFree(d_in); }
        if (cudaHostAlloc(&h_in, h2d_bytes, cudaHostAllocDefault) != cudaSuccess) return -2;
        if (cudaMalloc(&d_in, h2d_bytes) != cudaSuccess) return -2;
        cap_in = h2d_bytes;
    }
    if (cap_out < d2h_bytes) {
        if (h_out) { cudaFreeHost(h_out); cudaFree(d_out); }
        if (cudaHostAlloc(&h_out, d2h_bytes, cudaHostAllocDefault) != cudaSuccess) return -2;
        if (cudaMalloc(&d_out, d2h_bytes) != cudaSuccess) return -2;
        cap_out = d2h_bytes;
    }
    cudaDeviceSynchronize();
    cudaEventRecord(e0, s0);
    cudaStreamWaitEvent(s1, e0, 0);
    for (int r = 0; r < reps; ++r) {
        if (mode != 2) cudaMemcpyAsync(d_in, h_in, h2d_bytes, cudaMemcpyHostToDevice, s0);
        if (mode != 1) cudaMemcpyAsync(h_out, d_out, d2h_bytes, cudaMemcpyDeviceToHost, s1);
    }
    cudaEventRecord(j, s1);
    cudaStreamWaitEvent(s0, j, 0);
    cudaEventRecord(e1, s0);
    if (cudaEventSynchronize(e1) != cudaSuccess) return -2;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    *out_ms_per_rep = ms / reps;
    return 0;
}
