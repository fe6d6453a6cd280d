// Dependent-issue latency of the FP64 pipe (cycles per instruction on ONE dependent chain, one warp, nothing else on the SM):
// DFMA, DMUL, DADD, MUFU.RSQ64H + the 64-bit shuffle.  These set the floor of every latency-bound FP64 chain in the library
// (the 32-pivot chain of the Cholesky tile, the transcendental chains of the expression kernels at 6-8 warps per scheduler).
#include <cuda_runtime.h>

template <int OP>
__global__ void fp64_latency_kernel(double* out, long long* cycles, int iters, double seed)
{
    double x = seed + threadIdx.x * 1e-9, a = 1.0000001, b = 1e-9;
    const long long t0 = clock64();
    for (int k = 0; k < iters; ++k) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (OP == 0) x = fma(x, a, b);
            else if (OP == 1) x = x * a;
            else if (OP == 2) x = x + b;
            else if (OP == 3) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); x = y; }
            else x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) { cycles[0] = t1 - t0; out[0] = x; }
}

// out[5] = cycles per dependent DFMA, DMUL, DADD, MUFU.RSQ64H (as rsqrt.approx.ftz.f64), 64-bit SHFL
extern "C" int fadb_probe_fp64_latency(double* out)
{
    double* d = nullptr;
    long long* c = nullptr;
    if (cudaMalloc(&d, 8) != cudaSuccess || cudaMalloc(&c, 8) != cudaSuccess) return -2;
    const int iters = 4096;
    for (int op = 0; op < 5; ++op) {
        long long best = 1ll << 62;
        for (int rep = 0; rep < 4; ++rep) {
            switch (op) {
            case 0: fp64_latency_kernel<0><<<1, 32>>>(d, c, iters, 1.0); break;
            case 1: fp64_latency_kernel<1><<<1, 32>>>(d, c, iters, 1.0); break;
            case 2: fp64_latency_kernel<2><<<1, 32>>>(d, c, iters, 1.0); break;
            case 3: fp64_latency_kernel<3><<<1, 32>>>(d, c, iters, 1.5); break;
            default: fp64_latency_kernel<4><<<1, 32>>>(d, c, iters, 1.0); break;
            }
            long long h = 0;
            if (cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost) != cudaSuccess) return -3;
            if (h < best) best = h;
        }
        out[op] = (double)best / (16.0 * iters);
    }
    cudaFree(d); cudaFree(c);
    return 0;
}
