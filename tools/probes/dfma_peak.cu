// dfma_peak.cu — measured FP64 issue peak of the device ("confirm with a DFMA micro-benchmark").
// Every thread runs CHAINS independent register-resident DFMA chains; all SMs are filled with `ctas_per_sm` CTAs of 256 threads.
// Reports warp-level FP64 instructions per second (thread-instructions / s) and the equivalent TFLOP/s (2 flops per DFMA).
// Built by tools/probes/Makefile into tools/_build/libfadb_probes.so; run by tools/dfma_peak.py.  Not part of the product library.
#include <cuda_runtime.h>
#include <cstdio>

template <int CHAINS>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b)
{
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = 1.0 + 1e-9 * (threadIdx.x + c);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) x[c] = __fma_rn(x[c], a, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    if (s == 12345.678) out[0] = s;      // keeps the chains alive, never true
}

// out[0] = FP64 thread-instructions per second, out[1] = elapsed ms of the timed launch, out[2] = SM count
extern "C" int fadb_probe_dfma_peak(int ctas_per_sm, int chains, int iters, double* out)
{
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -2;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double* d = nullptr;
    if (cudaMalloc(&d, 8) != cudaSuccess) return -2;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = sms * ctas_per_sm;
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {          // first launches warm the clocks up; keep the best
        cudaEventRecord(e0);
        if (chains >= 8) dfma_kernel<8><<<grid, 256>>>(d, iters, 0.999999, 1e-6);
        else if (chains >= 4) dfma_kernel<4><<<grid, 256>>>(d, iters, 0.999999, 1e-6);
        else if (chains >= 2) dfma_kernel<2><<<grid, 256>>>(d, iters, 0.999999, 1e-6);
        else dfma_kernel<1><<<grid, 256>>>(d, iters, 0.999999, 1e-6);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) return -2;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const int ch = chains >= 8 ? 8 : chains >= 4 ? 4 : chains >= 2 ? 2 : 1;
    const double inst = (double)grid * 256.0 * (double)iters * 8.0 * ch;
    out[0] = inst / (best * 1e-3);
    out[1] = best;
    out[2] = sms;
    cudaFree(d);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}
