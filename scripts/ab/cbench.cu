// Micro-benchmark: FP64-pipe utilisation of the closed-form segment integrals (add_int_pn_tablelin<8>)
// as compiled, at 1..4 resident warps per SM sub-partition, no memory traffic.
#include <cstdio>
#include <cuda_runtime.h>
#include "../ndpp_b200/csrc/legendre.cuh"
using namespace ndpp;

template <int LT>
__global__ void __launch_bounds__(512, 1) k(double* out, int iters, double dx)
{
    double acc[NDPP_MAX_L];
#pragma unroll
    for (int l = 0; l < NDPP_MAX_L; ++l) acc[l] = 0.0;
    double x = -1.0 + 1e-4 * threadIdx.x, f = 0.3 + 1e-3 * threadIdx.x;
    for (int i = 0; i < iters; ++i) {
        const double xh = x + dx;
        const double fn = f * 1.0001;
        Powers A, B;
        make_powers(x, A);
        make_powers(xh, B);
        add_int_pn_tablelin<LT>(LT, x, xh, f, fn, A, B, acc);
        x = xh; f = fn;
    }
    double s = 0.0;
#pragma unroll
    for (int l = 0; l < NDPP_MAX_L; ++l) s += acc[l];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    double* out;
    cudaMalloc(&out, 148 * 512 * sizeof(double));
    const int iters = 20000;
    for (int w = 1; w <= 4; ++w) {
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        k<8><<<148, 128 * w>>>(out, 100, 1e-3);
        cudaEventRecord(a);
        k<8><<<148, 128 * w>>>(out, iters, 1e-5);
        cudaEventRecord(b);
        cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        const double cyc = ms * 1e-3 * 1.965e9;  // per SMSP
        printf("warps/SMSP %d: %.2f ms, %.1f cycles per segment per warp, %.1f cycles per segment per SMSP\n", w, ms,
               cyc / iters, cyc / iters / w);
    }
    return 0;
}
