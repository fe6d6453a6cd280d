// device_reduce.cuh — 256-bit global accesses and deterministic warp / block reductions (device only; also compiled by NVRTC
// for the specialised stage kernels, so no host or standard-library includes here).
#pragma once

namespace fadb {

// ---------------------------------------------------------------------------------------
// 256-bit global memory access (LDG.E.256 / STG.E.256 on sm_100a).
// Streaming data is read once: bypass L1 allocation; adjoint RMW uses plain ld/st.
// ---------------------------------------------------------------------------------------
struct __align__(32) double4_t { double v[4]; };

#ifndef FADB_NO_LD256
__device__ __forceinline__ double4_t ldg256_stream(const double* p)
{
    double4_t r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3]) : "l"(p));
    return r;
}
__device__ __forceinline__ double4_t ld256(const double* p)
{
    double4_t r;
    asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3]) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st256(double* p, const double4_t& r)
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "d"(r.v[0]), "d"(r.v[1]), "d"(r.v[2]), "d"(r.v[3]) : "memory");
}
#else   // assemblers older than PTX ISA 8.8 (NVRTC < 12.9): the same 32 bytes as two 128-bit accesses
__device__ __forceinline__ double4_t ldg256_stream(const double* p)
{
    double4_t r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.v[0]), "=d"(r.v[1]) : "l"(p));
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.v[2]), "=d"(r.v[3]) : "l"(p + 2));
    return r;
}
__device__ __forceinline__ double4_t ld256(const double* p)
{
    double4_t r;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.v[0]), "=d"(r.v[1]) : "l"(p) : "memory");
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.v[2]), "=d"(r.v[3]) : "l"(p + 2) : "memory");
    return r;
}
__device__ __forceinline__ void st256(double* p, const double4_t& r)
{
    asm volatile("st.global.v2.f64 [%0], {%1,%2};" :: "l"(p), "d"(r.v[0]), "d"(r.v[1]) : "memory");
    asm volatile("st.global.v2.f64 [%0], {%1,%2};" :: "l"(p + 2), "d"(r.v[2]), "d"(r.v[3]) : "memory");
}
#endif

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
