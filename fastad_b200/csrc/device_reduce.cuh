// device_reduce.cuh — deterministic warp / block reductions (device only; also compiled by NVRTC
// for the specialised stage kernels, so no host or standard-library includes here).
#pragma once

namespace fadb {

// Deterministic block reduction of NCH channels: warp shuffle tree, then shared memory,
// result valid in thread 0.  Fixed tree => run-to-run identical bits.
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_prod(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v *= __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

template <int NCH>
__device__ __forceinline__ void block_sum(double (&acc)[NCH], double* smem /* [32*NCH] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int c = 0; c < NCH; ++c) acc[c] = warp_sum(acc[c]);
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) smem[warp * NCH + c] = acc[c];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            double v = lane < nwarp ? smem[lane * NCH + c] : 0.0;
            acc[c] = warp_sum(v);
        }
    }
    __syncthreads();
}

}  // namespace fadb
