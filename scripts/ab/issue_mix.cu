// Does a non-FP64 instruction issue in the gap between two FP64 instructions (FP64 pipe: 16 lanes/clk/SMSP)?
#include <cstdio>
#include <cuda_runtime.h>
template <int NI>
__global__ void __launch_bounds__(512, 1) k(double* out, int* iout, int iters)
{
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    int i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3, i4 = i0 + 5, i5 = i0 + 7, i6 = i0 + 11, i7 = i0 + 13;
    const double m = 0.999999, b = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, b); if (NI > 0) i0 = i0 * 3 + i;
        a1 = fma(a1, m, b); if (NI > 1) i1 = i1 * 3 + i;
        a2 = fma(a2, m, b); if (NI > 2) i2 = i2 * 3 + i;
        a3 = fma(a3, m, b); if (NI > 3) i3 = i3 * 3 + i;
        a4 = fma(a4, m, b); if (NI > 4) i4 = i4 * 3 + i;
        a5 = fma(a5, m, b); if (NI > 5) i5 = i5 * 3 + i;
        a6 = fma(a6, m, b); if (NI > 6) i6 = i6 * 3 + i;
        a7 = fma(a7, m, b); if (NI > 7) i7 = i7 * 3 + i;
        if (NI > 8) { i0 ^= i1 >> 3; i2 ^= i3 >> 3; i4 ^= i5 >> 3; i6 ^= i7 >> 3; i1 += i2; i3 += i4; i5 += i6; i7 += i0; }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    iout[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = i0 + i1 + i2 + i3 + i4 + i5 + i6 + i7;
}
template <int NI> void run(double* out, int* iout)
{
    const int iters = 100000;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<NI><<<148, 512>>>(out, iout, 100);
    cudaEventRecord(a);
    k<NI><<<148, 512>>>(out, iout, iters);
    cudaEventRecord(b); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("8 DFMA + %d int per iteration, 4 warps/SMSP: %.3f ms, %.2f cycles per iteration per SMSP (per warp-iteration %.2f)\n", NI, ms,
           ms * 1e-3 * 1.965e9 / iters, ms * 1e-3 * 1.965e9 / iters / 4);
}
int main()
{
    double* out; int* iout;
    cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&iout, 148 * 512 * 4);
    run<0>(out, iout); run<4>(out, iout); run<8>(out, iout); run<9>(out, iout);
    return 0;
}
