// GPU: accuracy of include/fastad_b200/fastmath.cuh ON THE DEVICE, compiled -fmad=true exactly like
// fastad_b200/csrc/black_scholes.cu (MUFU seeds, contracted FMAs, the shared-memory copies of the tables), against glibc
// on the host.  Same argument ranges and bounds as the host-side tests/cpp/test_fastmath.cpp, over 2^21 points.
#include <cuda_runtime.h>

#include <random>
#include <vector>

#include <fastad_b200/fastmath.cuh>
#include "check.hpp"

constexpr int N = 1 << 21;
struct Out { double *expt, *logt, *logt1, *div, *sqrt, *rcp, *Phi, *PhiN, *E; };

__global__ void fm_kernel(int n, const double* x, const double* y, const double* y1, const double* p, const double* q,
                          const double* d, Out o)
{
    __shared__ double tab[675];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) tab[i] = fm::kErfcx32_d[i];
    if (threadIdx.x < 32) tab[256 + threadIdx.x] = fm::kExp2_32_d[threadIdx.x];
    for (int i = threadIdx.x; i < 387; i += blockDim.x) tab[288 + i] = fm::kLogT_d[i];
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        o.expt[i] = fm::exp_t(x[i], tab + 256);
        o.logt[i] = fm::log_t(y[i], tab + 288);
        o.logt1[i] = fm::log_t(y1[i], tab + 288);
        o.div[i] = fm::div_(p[i], q[i]);
        o.sqrt[i] = fm::sqrt_(p[i]);
        o.rcp[i] = fm::rcp_(q[i]);
        double Phi, PhiN, E;
        fm::phi_pair_pw(d[i], tab, Phi, PhiN, E);
        o.Phi[i] = Phi; o.PhiN[i] = PhiN; o.E[i] = E;
    }
}

#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

int main()
{
    std::mt19937_64 gen(12345);
    std::uniform_real_distribution<double> U(0., 1.);
    std::vector<double> x(N), y(N), y1(N), p(N), q(N), d(N);
    for (int i = 0; i < N; ++i) {
        const double a = U(gen), b = U(gen);
        x[i] = -60.0 * a * a + 2.0 * (b - 0.5);
        y[i] = std::exp(8.0 * (a - 0.5));
        y1[i] = 1.0 + (a - 0.5) * std::ldexp(1.0, -(int)(40 * b));
        p[i] = 0.01 + 300 * a; q[i] = 0.01 + 300 * b;
        d[i] = 16.0 * (a - 0.5);
    }
    // edge arguments the host test checks one by one
    x[0] = -800.0; x[1] = 0.0; x[2] = -708.0; x[3] = 709.0; x[4] = 1e-17;
    y[0] = 1.0; y[1] = 0.75; y[2] = 1.5; y[3] = 1e-300; y[4] = 1e300; y[5] = 2.2250738585072014e-308;
    d[0] = 0.0; d[1] = 40.0; d[2] = -1e-300; d[3] = -40.0;
    double* din[6];
    const std::vector<double>* hin[6] = {&x, &y, &y1, &p, &q, &d};
    for (int k = 0; k < 6; ++k) {
        CU(cudaMalloc(&din[k], N * sizeof(double)));
        CU(cudaMemcpy(din[k], hin[k]->data(), N * sizeof(double), cudaMemcpyHostToDevice));
    }
    double* dout[9];
    for (int k = 0; k < 9; ++k) CU(cudaMalloc(&dout[k], N * sizeof(double)));
    Out o{dout[0], dout[1], dout[2], dout[3], dout[4], dout[5], dout[6], dout[7], dout[8]};
    fm_kernel<<<148 * 4, 256>>>(N, din[0], din[1], din[2], din[3], din[4], din[5], o);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    std::vector<std::vector<double>> h(9, std::vector<double>(N));
    for (int k = 0; k < 9; ++k) CU(cudaMemcpy(h[k].data(), dout[k], N * sizeof(double), cudaMemcpyDeviceToHost));

    double e_expt = 0, e_logt = 0, e_logt1 = 0, e_div = 0, e_sqrt = 0, e_rcp = 0, e_abs = 0, e_tail = 0, e_E = 0;
    for (int i = 0; i < N; ++i) {
        if (x[i] >= -708.0) e_expt = std::max(e_expt, ulp_diff(h[0][i], std::exp(x[i])));
        e_logt = std::max(e_logt, ulp_diff(h[1][i], std::log(y[i])));
        e_logt1 = std::max(e_logt1, ulp_diff(h[2][i], std::log(y1[i])));
        e_div = std::max(e_div, ulp_diff(h[3][i], p[i] / q[i]));
        e_sqrt = std::max(e_sqrt, ulp_diff(h[4][i], std::sqrt(p[i])));
        e_rcp = std::max(e_rcp, ulp_diff(h[5][i], 1.0 / q[i]));
        const double u = d[i] * 0.70710678118654752440;
        const double ref = 0.5 * std::erfc(-u), refn = 0.5 * std::erfc(u);
        e_abs = std::max(e_abs, std::max(std::fabs(h[6][i] - ref), std::fabs(h[7][i] - refn)));
        if (std::fabs(d[i]) <= 8.0) e_tail = std::max(e_tail, d[i] < 0 ? ulp_diff(h[6][i], ref) : ulp_diff(h[7][i], refn));
        if (u * u < 708.0) e_E = std::max(e_E, ulp_diff(h[8][i], std::exp(-(u * u))));
    }
    std::printf("device (-fmad=true) max ulp over %d points: exp_t %.2f  log_t %.2f  log_t near 1 %.2f  div_ %.2f  sqrt_ %.2f  rcp_ %.2f  "
                "gauss %.2f | Phi abs %.3g, lower-tail ulp %.2f\n", N, e_expt, e_logt, e_logt1, e_div, e_sqrt, e_rcp, e_E, e_abs, e_tail);
    CHECK(e_expt <= 1.0);
    CHECK(e_logt <= 1.0);
    CHECK(e_logt1 <= 1.0);
    CHECK(e_div <= 1.0);
    CHECK(e_sqrt <= 1.0);
    CHECK(e_rcp <= 1.0);
    CHECK(e_E <= 1.0);
    CHECK(e_abs <= 4e-16);
    CHECK(e_tail <= 8.0);
    // edge cases, same as the host test
    CHECK(h[0][0] == 0.0 && h[0][1] == 1.0);                       // exp_t(-800) flushes, exp_t(0) == 1
    CHECK(h[1][0] == 0.0);                                         // log_t(1) == 0 exactly
    CHECK(std::fabs(h[6][0] - 0.5) <= 1.2e-16 && std::fabs(h[7][0] - 0.5) <= 1.2e-16 && h[8][0] == 1.0);
    CHECK(h[6][1] == 1.0 && h[7][1] == 0.0 && h[8][1] == 0.0);     // d = 40
    CHECK(h[6][3] == 0.0 && h[7][3] == 1.0);                       // d = -40
    CHECK(std::fabs(h[6][2] - 0.5) <= 1.2e-16);
    return report("test_fastmath_device");
}
