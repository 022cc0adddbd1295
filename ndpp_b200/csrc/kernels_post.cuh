// kernels_post.cuh -- the two steps that follow the integrator in the reference's driver
// (src/ndpp.F90:611-648), done on the device so that only the thinned matrices leave the GPU:
//
//   k_apply_tol   <- apply_tol_scatt  (src/scatt.F90:786-818)   one warp per E_in column
//   k_thin_grid   <- thin_grid_one / thin_grid_two (src/thin.F90:51-320)   one block, speculative
//   k_gather_cols    compaction of the kept columns
//
// thin_grid is a greedy scan: whether point k can be dropped depends on the last point kept (klo),
// so the decisions are sequential.  The block tests 32 consecutive candidates at once, one per warp,
// all against the current klo; the first candidate that must be kept becomes the new klo and the
// candidates after it are re-tested in the next round.  Decisions and kept values are those of the
// serial loop.  Of the two diagnostics only `compression` is reproduced exactly; the reference's
// `maxerr` compares a relative error with a stored absolute one in loop order (src/thin.F90:127-131)
// and is replaced by the plain maximum of |interpolated - y| over the accepted tests.
#pragma once
#include "common.cuh"

namespace ndpp {

__global__ void k_apply_tol(double* __restrict__ data, int NE, int G, int L, double tol)
{
    const int w = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= NE) return;
    double* col = data + (size_t)w * G * L;
    double orig_total = 0.0;
    for (int g = 0; g < G; ++g) orig_total = orig_total + col[g * L];   // same order on every lane
    __syncwarp();
    for (int g = lane; g < G; g += 32) {
        const double p0 = col[g * L];
        if ((p0 > 0.0) && (p0 < tol))
            for (int l = 0; l < L; ++l) col[g * L + l] = 0.0;
    }
    __syncwarp();
    double norm = 0.0;
    if (orig_total > 0.0) {
        double now = 0.0;
        for (int g = 0; g < G; ++g) now = now + col[g * L];
        norm = orig_total / now;
    }
    __syncwarp();
    for (int e = lane; e < G * L; e += 32) col[e] = col[e] * norm;
}

#define THIN_WARPS 32
// keep[] <- 0-based indices of the kept points; out[0] = number kept, maxabs[0] = max |interpolated - y|
__global__ void __launch_bounds__(THIN_WARPS * 32)
k_thin_grid(const double* __restrict__ x, const double* __restrict__ y1, const double* __restrict__ y2, int NE, int GL,
            const double* __restrict__ tokeep, int n_tokeep, double tol, int* __restrict__ keep, int* __restrict__ out,
            double* __restrict__ maxabs)
{
    __shared__ int s_fail[THIN_WARPS];
    __shared__ double s_max[THIN_WARPS];
    __shared__ int s_klo, s_base, s_nkeep;
    __shared__ double s_mabs;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        s_klo = 0; s_base = 1; s_nkeep = 0; s_mabs = 0.0;
        if (NE >= 1) keep[s_nkeep++] = 0;
    }
    __syncthreads();
    while (s_base <= NE - 2) {
        const int klo = s_klo, k = s_base + warp;
        int fail = 0;
        double mabs = 0.0;
        if (k <= NE - 2) {
            const double x1 = x[klo], x2 = x[k + 1], xx = x[k];
            const double x_frac = 1.0 / log(x2 / x1) * log(xx / x1);   // log interpolation (src/thin.F90:103)
            for (int t = 0; t < n_tokeep; ++t)
                if (tokeep[t] == xx) fail = 1;
            if (!fail) {
                const int nmat = y2 ? 2 : 1;
                for (int m = 0; m < nmat; ++m) {
                    const double* __restrict__ Y = m ? y2 : y1;
                    for (int e = lane; e < GL; e += 32) {
                        const double a = Y[(size_t)klo * GL + e], b = Y[(size_t)(k + 1) * GL + e], y = Y[(size_t)k * GL + e];
                        const double testval = a + (b - a) * x_frac;
                        double error = fabs(testval - y);
                        if (y != 0.0) error = error / y;   // signed, as the reference writes it
                        if (error <= tol) mabs = fmax(mabs, fabs(testval - y));
                        else fail = 1;
                    }
                }
                fail = __any_sync(0xffffffffu, fail);
                for (int o = 16; o > 0; o >>= 1) mabs = fmax(mabs, __shfl_xor_sync(0xffffffffu, mabs, o));
            }
        }
        if (lane == 0) { s_fail[warp] = (k <= NE - 2) ? fail : -1; s_max[warp] = mabs; }
        __syncthreads();
        if (threadIdx.x == 0) {
            int f = -1;
            for (int w = 0; w < THIN_WARPS; ++w) {
                if (s_fail[w] < 0) break;
                s_mabs = fmax(s_mabs, s_max[w]);   // tests of the candidates up to and including the first kept one
                if (s_fail[w]) { f = w; break; }
            }
            if (f >= 0) { keep[s_nkeep++] = s_base + f; s_klo = s_base + f; s_base = s_base + f + 1; }
            else s_base = s_base + THIN_WARPS;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (NE > 1) keep[s_nkeep++] = NE - 1;
        out[0] = s_nkeep;
        maxabs[0] = s_mabs;
    }
}

// dst[j][:] = src[keep[j]][:], j < n_keep (n_keep read from the device); src and dst must not overlap
__global__ void k_gather_cols(const double* __restrict__ src, const int* __restrict__ keep, const int* __restrict__ n_keep,
                              int width, double* __restrict__ dst)
{
    const int n = *n_keep;
    for (int j = blockIdx.x; j < n; j += gridDim.x) {
        const double* s = src + (size_t)keep[j] * width;
        double* d = dst + (size_t)j * width;
        for (int e = threadIdx.x; e < width; e += blockDim.x) d[e] = s[e];
    }
}

}  // namespace ndpp
