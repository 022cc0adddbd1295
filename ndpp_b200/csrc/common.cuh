// common.cuh -- device-side data layout and small helpers shared by all kernels.
//
// Data layout in HBM (structure of arrays, one set per ScattData slot):
//   tab   [sum_i NP_i][M]  uniform-mu tables, one contiguous M-row per (E_in row i, E_out point j):
//                          tab[(row_off[i] + j) * M + k] == Fortran distro(i)%data(k+1, j+1)
//   eout / pdf / cdf [sum_i NP_i]   outgoing-energy grids of the rows, same row_off offsets
//   e_grid[NE], intt[NE], row_off[NE+1]
// Indices inside device code are 0-based; every formula that the reference writes with 1-based
// indices is shifted explicitly so that the floating-point operands are identical.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "legendre.cuh"

namespace ndpp {

constexpr double FP_PRECISION = 1e-14;          // src/constants.F90:22
constexpr double REF_PI = 3.1415926535898;      // src/constants.F90:35 (truncated on purpose)
constexpr double REF_INFINITY = 1.7976931348623157e308;
constexpr double MIN_EIN = 1e-14;               // src/constants.F90:109
enum { HISTOGRAM = 1, LINEAR_LINEAR = 2, LINEAR_LOG = 3, LOG_LINEAR = 4, LOG_LOG = 5 };
enum { ANGLE_ISOTROPIC = 1, ANGLE_32_EQUI = 2, ANGLE_TABULAR = 3 };
enum { MT_ELASTIC = 2 };

// Device view of one ScattData slot (src/scattdata_header.F90:36-69).
struct SlotDev {
    int NE;                 // incoming energies of the distribution
    int M;                  // mu_bins
    int law;                // 0, 3, 4, 9, 44, 61
    int has_adist, has_edist, scatter_in_cm, MT;
    int threshold;          // 1-based index into the nuclide grid
    int n_sigma;
    int multiplicity;
    int total_np;           // sum_i NP_i
    int max_np;             // max_i NP_i
    double Q;
    const double* e_grid;   // [NE]
    const int* row_off;     // [NE+1]
    const int* intt;        // [NE]
    double* tab;            // [total_np][M]
    const double* eout;     // [total_np]
    const double* pdf;
    const double* cdf;
    const double* sigma;    // [n_sigma] (elastic: the nuclide's elastic xs)
    const double* p_valid;  // flattened TAB1 or nullptr
    const double* yield;    // flattened TAB1 or nullptr
    // raw ACE blocks (inputs of convert_distro)
    const double* ad_energy; const int* ad_type; const int* ad_loc; const double* ad_data; int ad_n;
    const double* ed_data;  int edist_law;
};

// Device view of the nuclide-level data (src/ace_header.F90:94-112 + group structure).
struct NucDev {
    int n_grid, n_bins, M, L, G;
    int ne_per_grp, adaptive_mu_its, adaptive_eout_its;
    double awr, kT, freegas_cutoff;
    double sab_threshold, brent_mu_thresh, adaptive_mu_tol, adaptive_eout_tol;
    const double* energy;   // [n_grid]
    const double* elastic;  // [n_grid]
    const double* e_bins;   // [n_bins]
    const double* mu;       // [M]
    int* err;               // device error word: 1 = binary_search out of range (src/search.F90:36-38)
};

// binary_search_real, src/search.F90:21-71: 0-based lower index i with a[i] <= val < a[i+1]
// (n-2 for val == a[n-1]).  The caller guarantees a[0] <= val <= a[n-1]; out-of-range values are
// clamped (the reference aborts) and flagged through *err when err != nullptr.
__device__ __forceinline__ int binary_search(const double* __restrict__ a, int n, double val, int* err = nullptr)
{
    int L = 0, R = n - 1;
    if (val < a[L] || val > a[R]) {
        if (err) *err = 1;
        return (val < a[L]) ? 0 : (n > 1 ? n - 2 : 0);
    }
    int it = 0;
    while (R - L > 1) {
        if (val > a[L] && val < a[L + 1]) return L;
        if (val > a[R - 1] && val < a[R]) return R - 1;
        const int mid = L + (R - L) / 2;
        if (val >= a[mid]) L = mid; else R = mid;
        if (++it == 64) break;
    }
    return L;
}

// interpolate_tab1 on a flattened TAB1 [NR, NBT(NR), INT(NR), NP, x(NP), y(NP)],
// src/interpolation.F90:24-123.
__device__ __forceinline__ double interpolate_tab1(const double* __restrict__ d, double x)
{
    const int nr = (int)d[0];
    const double* nbt = d + 1;
    const double* itp = d + 1 + nr;
    const int np = (int)d[1 + 2 * nr];
    const double* xs = d + 2 + 2 * nr;
    const double* ys = xs + np;
    if (x < xs[0]) return ys[0];
    if (x > xs[np - 1]) return ys[np - 1];
    const int i = binary_search(xs, np, x);
    int interp = LINEAR_LINEAR;
    if (nr == 1) {
        interp = (int)itp[0];
    } else if (nr > 1) {
        for (int j = 0; j < nr; ++j)
            if ((double)(i + 1) < nbt[j]) { interp = (int)itp[j]; break; }
    }
    if (interp == HISTOGRAM) return ys[i];
    const double x0 = xs[i], x1 = xs[i + 1], y0 = ys[i], y1 = ys[i + 1];
    double r;
    switch (interp) {
    case LINEAR_LINEAR: r = (x - x0) / (x1 - x0); return (1 - r) * y0 + r * y1;
    case LINEAR_LOG: r = (log(x) - log(x0)) / (log(x1) - log(x0)); return (1 - r) * y0 + r * y1;
    case LOG_LINEAR: r = (x - x0) / (x1 - x0); return exp((1 - r) * log(y0) + r * log(y1));
    case LOG_LOG: r = (log(x) - log(x0)) / (log(x1) - log(x0)); return exp((1 - r) * log(y0) + r * log(y1));
    default: return 0.0;
    }
}

// tolab, src/scattdata_header.F90:1466-1496: lab cosine u of CM cosine w for reduced mass R.
__device__ __forceinline__ double tolab(double R, double w)
{
    if (R > 1.0) return (1.0 + R * w) / sqrt(1.0 + R * R + 2.0 * R * w);
    if (R == 1.0) {
        if (w == -1.0) return -1.0;
        return (1.0 + R * w) / sqrt(1.0 + R * R + 2.0 * R * w);
    }
    if (w < -R) {
        double u = sqrt(1.0 - R * R);
        const double f = (w - (-1.0)) / (-R - 1.0);
        u = (1.0 - f) * (-1.0) + f * u;
        return u;
    }
    return (1.0 + R * w) / sqrt(1.0 + R * R + 2.0 * R * w);
}

// The same search for a caller that guarantees 0 <= a[0] <= val <= a[n-1] (non-negative values
// order like their bit patterns, so the comparisons run on the integer pipe).  The reference's two
// early-exit tests only shortcut cases in which the bisection reaches the same index, so they are
// dropped; the bisection itself follows the reference's mid-point rule step for step.
__device__ __forceinline__ int binary_search_nonneg(const double* __restrict__ a, int n, double val)
{
    int L = 0, R = n - 1;
    const long long v = __double_as_longlong(val);
    while (R - L > 1) {
        const int mid = L + (R - L) / 2;
        if (v >= __double_as_longlong(a[mid])) L = mid; else R = mid;
    }
    return L;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Per-(E_in, slot) scalars of scatt_interp_distro, src/scattdata_header.F90:416-497.
struct InterpInfo {
    int active;      // 0 => the distribution is left at zero
    int iE;          // 0-based lower table row
    int scaled;      // non-elastic: result is multiplied by sigS, then by p_valid (:494-497)
    double sigS;     // sigma_s(E_in)
    double p_valid;  // probability of this (nested) law
};

__device__ __forceinline__ double apply_scale(const InterpInfo& r, double v)
{
    return r.scaled ? (v * r.sigS) * r.p_valid : v;
}

__device__ __forceinline__ InterpInfo interp_info(const NucDev& nuc, const SlotDev& s, double Ein)
{
    InterpInfo r;
    r.active = 0; r.iE = 0; r.sigS = 0.0; r.p_valid = 1.0; r.scaled = (s.MT != MT_ELASTIC);
    const double Ethr = nuc.energy[s.threshold - 1];
    if (((Ein <= Ethr) && (s.threshold > 1)) || (Ein > nuc.e_bins[nuc.n_bins - 1])) return r;
    if (Ein >= nuc.energy[nuc.n_grid - 1]) {
        r.sigS = s.sigma[s.n_sigma - 1];
        r.iE = s.NE - 2;                               // integrate_distro(this, Ein, NE - 1), 1-based
    } else {
        int k;                                         // 0-based nuclide grid index
        if (Ein <= nuc.energy[0]) k = 0; else k = binary_search(nuc.energy, nuc.n_grid, Ein);
        if (nuc.energy[k] == nuc.energy[k + 1]) k = k + 1;
        const double f = (Ein - nuc.energy[k]) / (nuc.energy[k + 1] - nuc.energy[k]);
        const int ks = k - (s.threshold - 1);          // index into sigma
        r.sigS = (1.0 - f) * s.sigma[ks] + f * s.sigma[ks + 1];
        if (r.sigS <= 0.0) return r;
        int iE;
        if (Ein < s.e_grid[0]) iE = 0; else iE = binary_search(s.e_grid, s.NE, Ein, nuc.err);
        if (s.e_grid[iE] >= s.e_grid[iE + 1]) iE = iE + 1;
        r.iE = iE;
    }
    if (s.has_edist) r.p_valid = interpolate_tab1(s.p_valid, Ein);
    r.active = 1;
    return r;
}

}  // namespace ndpp
