// kernels_convert.cuh -- K1: ACE angular / energy-angle laws -> uniform-mu tables, on the device.
//
// Replaces scatt_convert_distro / convert_file4 / convert_file6
// (src/scattdata_header.F90:325-382, 669-762, 769-950).  One thread per table element
// (row, E_out column, mu point).  The reference walks each table with a running `idata_prev`
// cursor; because the cursor only ever skips entries that already failed the match test for a
// smaller mu, a search from the start of the block finds the same entry, which is what makes the
// conversion data-parallel without changing any result.
#pragma once
#include "common.cuh"
#include "libm_exact.cuh"

namespace ndpp {

#define ADATA(k) data[(k)-1]  // Fortran 1-based data(k)

// convert_file4 for one mu point (src/scattdata_header.F90:686-752); 0 when nothing matches.
__device__ __forceinline__ double convert_file4_point(const SlotDev& s, int iEad /*0-based*/, double mu, int imu)
{
    const double* data = s.ad_data;
    int lc = s.ad_loc[iEad];
    switch (s.ad_type[iEad]) {
    case ANGLE_ISOTROPIC: return 0.5;
    case ANGLE_32_EQUI:
        for (int idata = lc + 1; idata <= lc + 1 + 32; ++idata) {
            if (ADATA(idata) >= mu) {
                if (imu == 0) return (1.0 / 32.0) / (ADATA(idata + 1) - ADATA(idata));
                return (1.0 / 32.0) / (ADATA(idata) - ADATA(idata - 1));
            }
        }
        return 0.0;
    case ANGLE_TABULAR: {
        const int interp = (int)ADATA(lc + 1);
        const int NP = (int)ADATA(lc + 2);
        lc = lc + 3;
        if (interp == HISTOGRAM) {
            for (int idata = lc; idata <= lc + NP - 1; ++idata) {
                if ((ADATA(idata) - mu) > FP_PRECISION) return ADATA(idata - 1 + NP);
                if (fabs(ADATA(idata) - mu) <= FP_PRECISION) return ADATA(idata + NP);
            }
        } else if (interp == LINEAR_LINEAR) {
            for (int idata = lc; idata <= lc + NP - 1; ++idata) {
                if ((ADATA(idata) - mu) > FP_PRECISION) {
                    const double r = (mu - ADATA(idata - 1)) / (ADATA(idata) - ADATA(idata - 1));
                    return ADATA(idata + NP - 1) + r * (ADATA(idata + NP) - ADATA(idata + NP - 1));
                }
                if (fabs(ADATA(idata) - mu) <= FP_PRECISION) return ADATA(idata + NP);
            }
        }
        return 0.0;
    }
    default: return 0.0;
    }
}
#undef ADATA

// Slots with law 0 / 3 / 9: one column per row (src/scattdata_header.F90:345-350).
__global__ void k_convert_file4(SlotDev s, const double* __restrict__ mu)
{
    const int M = s.M;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)s.NE * M) return;
    const int iE = (int)(t / M), imu = (int)(t % M);
    // A law-9 slot takes NE from its energy distribution but indexes the angular distribution with the same
    // row number (:345-350); past the end of adist the reference reads out of bounds (undefined).  Rows beyond
    // the last angular entry use that entry (the synthesised adist is isotropic throughout).
    const int iEad = iE < s.ad_n ? iE : s.ad_n - 1;
    s.tab[(size_t)s.row_off[iE] * M + imu] = convert_file4_point(s, iEad, mu[imu], imu);
}

#define EDATA(k) data[(k)-1]

// One Law-61 angular table evaluated at mu (src/scattdata_header.F90:843-945).
__device__ __forceinline__ double law61_point(const double* data, int lc, double mu)
{
    const int interp = (int)EDATA(lc + 1);
    const int NPang = (int)EDATA(lc + 2);
    lc = lc + 3;
    // an unknown interpolation code is the reference's fatal_error (:944); the host rejects it in
    // ndppgpu_nuclide_add_reaction before any table is built (validate_law61)
    if (interp < HISTOGRAM || interp > LOG_LOG) return 0.0;
    for (int idata = lc; idata <= lc + NPang - 1; ++idata) {
        if ((EDATA(idata) - mu) > FP_PRECISION) {
            double r;
            switch (interp) {
            case HISTOGRAM: return EDATA(idata + NPang - 1);
            case LINEAR_LINEAR:
                r = (mu - EDATA(idata - 1)) / (EDATA(idata) - EDATA(idata - 1));
                return EDATA(idata + NPang - 1) + r * (EDATA(idata + NPang) - EDATA(idata - 1 + NPang));
            case LINEAR_LOG:
                r = (log(mu) - log(EDATA(idata - 1))) / (log(EDATA(idata)) - log(EDATA(idata - 1)));
                return EDATA(idata + NPang - 1) + r * (EDATA(idata + NPang) - EDATA(idata - 1 + NPang));
            case LOG_LINEAR:
                r = (mu - EDATA(idata - 1)) / (EDATA(idata) - EDATA(idata - 1));
                return lm::exp_((1.0 - r) * log(EDATA(idata + NPang)) + r * log(EDATA(idata + NPang - 1)));
            default:  // LOG_LOG
                r = (log(mu) - log(EDATA(idata - 1))) / (log(EDATA(idata)) - log(EDATA(idata - 1)));
                return lm::exp_((1.0 - r) * log(EDATA(idata + NPang)) + r * log(EDATA(idata + NPang - 1)));
            }
        }
        if (fabs(EDATA(idata) - mu) <= FP_PRECISION) return EDATA(idata + NPang);
    }
    return 0.0;
}

// Slots with law 4 / 44 / 61: one thread per (column, mu).  Thread imu == 0 of a column also
// copies E_out / pdf / cdf (src/scattdata_header.F90:806-817); eout/pdf/cdf/intt are written
// through non-const aliases held by the host.
__global__ void k_convert_file6(SlotDev s, const double* __restrict__ mu, double* __restrict__ eout,
                                double* __restrict__ pdf, double* __restrict__ cdf, int* __restrict__ intt)
{
    const int M = s.M;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)s.total_np * M) return;
    const int col = (int)(t / M), imu = (int)(t % M);
    // row of this column
    int lo = 0, hi = s.NE;  // row_off[lo] <= col < row_off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s.row_off[mid] <= col) lo = mid; else hi = mid;
    }
    const int iE = lo, iEout = col - s.row_off[iE];  // 0-based
    const double* data = s.ed_data;
    const int NR = (int)EDATA(1);                    // NR > 0 is rejected on the host (:797-800)
    const int NE = (int)EDATA(2 + 2 * NR);
    const int lc0 = (int)EDATA(2 + 2 * NR + NE + (iE + 1));
    const int NP = (int)EDATA(lc0 + 2);
    if (imu == 0) {
        eout[col] = EDATA(lc0 + 2 + (iEout + 1));
        pdf[col] = EDATA(lc0 + 2 + NP + (iEout + 1));
        cdf[col] = EDATA(lc0 + 2 + 2 * NP + (iEout + 1));
        if (iEout == 0) {
            int it = (int)EDATA(lc0 + 1);
            if (it > 10) it = it % 10;
            intt[iE] = it;
        }
    }
    double v = 0.0;
    const double x = mu[imu];
    if (s.edist_law == 44) {
        const int lc = lc0 + 2;
        const double KMR = EDATA(lc + 3 * NP + (iEout + 1));
        const double KMA = EDATA(lc + 4 * NP + (iEout + 1));
        // sinh / cosh with the bits of the host libm the reference calls (libm_exact.cuh)
        const double KMconst = 0.5 * KMA / lm::sinh_(KMA);
        v = KMconst * (lm::cosh_(KMA * x) + KMR * lm::sinh_(KMA * x));
    } else if (s.edist_law == 61) {
        const int lcin = lc0 + 2;
        const int lc = (int)EDATA(lcin + 3 * NP + (iEout + 1));
        v = (lc == 0) ? 0.5 : law61_point(data, lc, x);
    } else {  // law 4 with an angular distribution (:351-372)
        // convert_file4 fills column 1; the copy loop runs over size(Eouts) == 2 columns only
        // (Eouts was just allocated with 2 entries by convert_file4), the rest stays zero.
        if (iEout < 2) {
            const double Eg = s.e_grid[iE];
            int iEad;
            if (Eg <= s.ad_energy[0]) iEad = 0;
            else if (Eg >= s.ad_energy[s.ad_n - 1]) iEad = s.ad_n - 1;
            else iEad = binary_search(s.ad_energy, s.ad_n, Eg);
            v = convert_file4_point(s, iEad, x, imu);
        }
    }
    s.tab[(size_t)col * M + imu] = v;
}
#undef EDATA

}  // namespace ndpp
