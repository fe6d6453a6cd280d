// forward_batch.cuh (nvcc only) — batched forward-mode sweeps on the GPU.
//
//   ad::forward_batch<NIN>(n, f, in_val, in_tan, out_val, out_tan [, stream]);
//
// evaluates, for every instance i < n, y = f(x_0 ... x_{NIN-1}) with x_k = ForwardVar(in_val[k][i],
// in_tan[k] ? in_tan[k][i] : dir[k]) and stores y.get_value() / y.get_adjoint().  One instance per
// thread: the whole dual-number computation stays in registers (no tape, no graph), inputs are
// SoA device arrays read coalesced, the grid is one resident wave with a grid-stride loop.  `f`
// is any __device__ callable taking NIN ForwardVar<double> and returning one, e.g. a
// __device__ lambda (compile with --extended-lambda) or a functor struct.
// This is the forward-mode counterpart of the batched reverse-mode kernels: the reference has
// only the scalar host type (forward/core/forward.hpp); the batch loop around it is the user's.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <utility>

#include "../fastad_b200.h"
#include "dual.hpp"

namespace ad {
namespace detail {

template <int NIN>
struct FwdBatchArgs {
    const double* val[NIN];
    const double* tan[NIN];     // nullptr: constant direction dir[k] for every instance
    double dir[NIN];
    double* out_val;
    double* out_tan;
};

template <class F, int NIN, size_t... I>
__device__ __forceinline__ ForwardVar<double> fwd_apply(const F& f, const ForwardVar<double> (&x)[NIN], std::index_sequence<I...>)
{
    return f(x[I]...);
}

template <class F, int NIN>
__global__ void __launch_bounds__(256) forward_batch_kernel(size_t n, F f, FwdBatchArgs<NIN> a)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        ForwardVar<double> x[NIN];
#pragma unroll
        for (int k = 0; k < NIN; ++k) x[k] = ForwardVar<double>(a.val[k][i], a.tan[k] ? a.tan[k][i] : a.dir[k]);
        const ForwardVar<double> y = fwd_apply<F, NIN>(f, x, std::make_index_sequence<NIN>{});
        if (a.out_val) a.out_val[i] = y.get_value();
        if (a.out_tan) a.out_tan[i] = y.get_adjoint();
    }
}

}  // namespace detail

// in_val[k]: device array of n values of input k; in_tan[k]: device array of n tangents or nullptr
// (then dir[k] is the tangent of input k for every instance).  Returns a cudaError_t as int.
template <int NIN, class F>
inline int forward_batch(size_t n, F f, const double* const (&in_val)[NIN], const double* const (&in_tan)[NIN],
                         const double (&dir)[NIN], double* out_val, double* out_tan, cudaStream_t stream = nullptr)
{
    if (n == 0) return 0;
    detail::FwdBatchArgs<NIN> a;
    for (int k = 0; k < NIN; ++k) { a.val[k] = in_val[k]; a.tan[k] = in_tan[k]; a.dir[k] = dir[k]; }
    a.out_val = out_val;
    a.out_tan = out_tan;
    if (!stream) {
        void* s = nullptr;
        if (fadb_get_stream(&s) != 0) return (int)cudaErrorUnknown;
        stream = reinterpret_cast<cudaStream_t>(s);
    }
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, detail::forward_batch_kernel<F, NIN>, 256, 0);
    size_t grid = (n + 255) / 256;
    const size_t wave = (size_t)sms * (per_sm > 0 ? per_sm : 1);
    if (grid > wave) grid = wave;
    detail::forward_batch_kernel<F, NIN><<<(unsigned)grid, 256, 0, stream>>>(n, f, a);
    return (int)cudaGetLastError();
}

}  // namespace ad
