// kernels_chi.cuh -- fission-spectrum (chi) integration, SURVEY 8f row N4.
//
// calc_chi's E_in loop (src/chi.F90:120-153) with chi_integrate / chi_prob / chi_beta
// (src/chidata_header.F90:143-494) and nu_total / nu_delayed (src/fission.F90:18-103).
// One block per incoming energy: thread s integrates law s over the groups (the law-4 walk through the
// cdf and every normalising sum are sequential in the reference and stay sequential in one thread, so
// the additions happen in the reference's order), then the threads turn to the groups and combine the
// laws in the reference's order.  The arithmetic is a few hundred exp/erf per E_in -- this kernel is
// latency-bound and tiny; it is on the device so that a nuclide's whole library record is produced
// without a host detour, not because it needs the FP64 pipe.
//
// Reference behaviour kept on purpose (each item has a test in tests/test_chi.py):
//  - law 7 replaces a group edge above Ein-U by U (:358,362); laws 7/9/11 return zeros without the final
//    normalisation when Ein <= U; laws that only warn (1,3,5,12,44,66,67) reach it and become NaN (:483-492);
//  - law 4/61 reads the cdf lin-lin whatever INTT' says (:313-320);
//  - chi_total = chi_prompt * (1 + prob of the LAST prompt law) before the delayed terms (src/chi.F90:131).
#pragma once
#include "common.cuh"

namespace ndpp {

struct ChiSlotDev {   // == ndppgpu_chi_slot (include/ndppgpu.h)
    int law, delayed, precursor, threshold, use_pvalid, n_sigma, sigma_off, data_off, pvalid_off, reserved;
};

struct ChiDev {
    int n_grid, n_bins, n_slots, n_prompt, n_precursor, nu_t_type, nu_d_type, NE;
    const double* energy;     // [n_grid]
    const double* fission;    // [n_grid]
    const double* nu_t_data;
    const double* nu_d_data;
    const double* precursor;  // nu_d_precursor_data
    const ChiSlotDev* slots;
    const double* pool;
    const double* e_bins;     // [n_bins]
    const double* Ein;        // [NE]
    double* work;             // [NE][n_prompt][G] prompt spectra
    double* chi_total;        // [NE][G]
    double* chi_prompt;       // [NE][G]
    double* chi_delay;        // [n_precursor][NE][G]
    int* err;                 // 1 = binary search out of range, 2 = NR > 1, 3 = discrete lines, 4 = no nu data
};

// libgcc's __powidf2, which gfortran emits for E**i with an integer variable exponent
__device__ __forceinline__ double powi_ref(double x, int m)
{
    unsigned n = (m < 0) ? (unsigned)(-m) : (unsigned)m;
    double y = (n % 2) ? x : 1.0;
    while (n >>= 1) {
        x = x * x;
        if (n % 2) y = y * x;
    }
    return (m < 0) ? 1.0 / y : y;
}

__device__ inline double chi_nu_total(const ChiDev& c, double E)
{
    if (c.nu_t_type == 1) {
        const int NC = (int)c.nu_t_data[0];
        double nu = 0.0;
        for (int i = 0; i <= NC - 1; ++i) nu = nu + c.nu_t_data[i + 1] * powi_ref(E, i);
        return nu;
    }
    if (c.nu_t_type == 2) return interpolate_tab1(c.nu_t_data, E);
    *c.err = 4;
    return 0.0;
}

__device__ inline double chi_prob(const ChiDev& c, const ChiSlotDev& s, double Ein)
{
    if (s.delayed) {
        int lc = 0;   // 0-based index of the group's decay constant
        for (int j = 1; j <= c.n_precursor; ++j) {
            const int NR = (int)c.precursor[lc + 1];
            const int NE = (int)c.precursor[lc + 2 + 2 * NR];
            if (j == s.precursor) break;
            lc = lc + 2 + 2 * NR + 2 * NE + 1;
        }
        return interpolate_tab1(c.precursor + lc + 1, Ein);
    }
    int j;   // 0-based lower grid index
    double f;
    if (Ein < c.energy[0]) {
        j = 0; f = 0.0;
    } else if (Ein >= c.energy[c.n_grid - 1]) {
        j = c.n_grid - 2; f = 1.0;
    } else {
        j = binary_search(c.energy, c.n_grid, Ein);
        f = (Ein - c.energy[j]) / (c.energy[j + 1] - c.energy[j]);
    }
    if (c.energy[j] == c.energy[j + 1]) j = j + 1;
    double prob;
    if (j + 1 < s.threshold) {
        prob = 0.0;
    } else {
        const double* sigma = c.pool + s.sigma_off;
        const int k = j + 1 - s.threshold;   // 0-based index of sigma(j - threshold + 1)
        prob = ((1.0 - f) * sigma[k] + f * sigma[k + 1]) / ((1.0 - f) * c.fission[j] + f * c.fission[j + 1]);
    }
    if (s.use_pvalid) prob = prob * interpolate_tab1(c.pool + s.pvalid_off, Ein);
    return prob;
}

// chi_integrate: chis[0..G) for one law at one E_in.  `d1` is the law data shifted so that d1[i] is the
// reference's edist % data(i) (1-based), which keeps the locator arithmetic of the Fortran text.
__device__ inline void chi_integrate(const ChiDev& c, const ChiSlotDev& s, double Ein, double* __restrict__ chis)
{
    const int G = c.n_bins - 1;
    const double* d1 = c.pool + s.data_off - 1;
    const double* Eb1 = c.e_bins - 1;   // Eb1[g] = E_bins(g)
    for (int g = 0; g < G; ++g) chis[g] = 0.0;
    switch (s.law) {
    case 4:
    case 61: {
        bool histogram_interp = false;
        const int NR = (int)d1[1];
        const int NE = (int)d1[2 + 2 * NR];
        if (NR == 1) {
            if (s.law == 4) histogram_interp = (d1[3] == 1.0);
        } else if (NR > 1) {
            *c.err = 2;
            return;
        }
        int lc = 2 + 2 * NR, iE;
        double x;
        if (Ein < d1[lc + 1]) {
            iE = 1; x = 0.0;
        } else if (Ein >= d1[lc + NE]) {
            iE = NE - 1; x = 1.0;
        } else {
            iE = binary_search(d1 + lc + 1, NE, Ein) + 1;
            x = (Ein - d1[lc + iE]) / (d1[lc + iE + 1] - d1[lc + iE]);
        }
        if (!histogram_interp && x > 0.5) iE = iE + 1;
        lc = (int)d1[2 + 2 * NR + NE + iE];
        const int INTTp = (int)d1[lc + 1];
        const int NP = (int)d1[lc + 2];
        if (INTTp > 10 && (INTTp - INTTp % 10) / 10 > 0) {
            *c.err = 3;
            return;
        }
        lc = lc + 3;
        int lEout_min = lc;
        double runsum = 0.0;
        for (int g = 1; g <= G; ++g) {
            const double top = Eb1[g + 1];
            for (iE = lEout_min; iE <= NP + lc - 2; ++iE)
                if (d1[iE + 1] > top) break;
            if (iE == NP + lc - 1) iE = iE - 1;
            const double interp = (top - d1[iE]) / (d1[iE + 1] - d1[iE]);
            double v = (d1[iE + 2 * NP] + interp * (d1[iE + 1 + 2 * NP] - d1[iE + 2 * NP]));
            v = v - runsum;
            runsum = runsum + v;
            chis[g - 1] = v;
            lEout_min = iE;
        }
        break;
    }
    case 7: {
        const int NR = (int)d1[1];
        const int NE = (int)d1[2 + 2 * NR];
        const double T = interpolate_tab1(d1 + 1, Ein);
        const double U = d1[2 + 2 * NR + 2 * NE + 1];
        if (Ein - U <= 0.0) return;
        const double x = (Ein - U) / T;
        const double I = sqrt(T * T * T) * (sqrt(0.25 * REF_PI) * erf(x) - x * exp(-x));
        for (int g = 1; g <= G; ++g) {
            double Egp1 = Eb1[g + 1];
            if (Egp1 > Ein - U) Egp1 = U;
            double v = 0.5 * (sqrt(REF_PI * T) * erf(sqrt(Egp1 / T)) * exp(Egp1 / T) - 2.0 * sqrt(Egp1)) * T * exp(-Egp1 / T);
            double Eg = Eb1[g];
            if (Eg > Ein - U) Eg = U;
            v = v - (0.5 * (sqrt(REF_PI * T) * erf(sqrt(Eg / T)) * exp(Eg / T) - 2.0 * sqrt(Eg)) * T * exp(-Eg / T));
            chis[g - 1] = v / I;
        }
        break;
    }
    case 9: {
        const int NR = (int)d1[1];
        const int NE = (int)d1[2 + 2 * NR];
        const double T = interpolate_tab1(d1 + 1, Ein);
        const double U = d1[2 + 2 * NR + 2 * NE + 1];
        const double x = (Ein - U) / T;
        if (Ein - U <= 0.0) return;
        const double ex = exp(x);
        for (int g = 1; g <= G; ++g) {
            double Egp1 = Eb1[g + 1], Eg = Eb1[g];
            if (Egp1 > (Ein - U)) Egp1 = Ein - U;
            if (Eg > (Ein - U)) Eg = Ein - U;
            double v = (Egp1 * ex + T * ex) * exp(-Egp1 / T);
            v = v - (Eg * ex + T * ex) * exp(-Eg / T);
            chis[g - 1] = v / (T * (x - ex + 1.0));
        }
        break;
    }
    case 11: {
        int NR = (int)d1[1];
        int NE = (int)d1[2 + 2 * NR];
        const double Watt_a = interpolate_tab1(d1 + 1, Ein);
        int lc = 2 + 2 * (NR + NE);
        double Watt_b = interpolate_tab1(d1 + lc + 1, Ein);
        NR = (int)d1[lc + 1];
        NE = (int)d1[lc + 2 + 2 * NR];
        lc = lc + 2 + 2 * (NR + NE);
        const double U = d1[lc + 1];
        double x = (Ein - U) / Watt_a;
        if (Ein - U <= 0.0) return;
        const double x0 = Watt_a * Watt_b * 0.25;
        const double I = 0.25 * sqrt(REF_PI * (Watt_a * Watt_a * Watt_a) * Watt_b) * exp(x0) *
                             (erf(sqrt(x) - sqrt(x0)) + erf(sqrt(x) + sqrt(x0))) -
                         Watt_a * exp(-x * sinh(Watt_a * Watt_b * x));
        Watt_b = sqrt(Watt_b);
        x = sqrt(REF_PI * Watt_a) * Watt_b * exp(0.25 * Watt_a * (Watt_b * Watt_b));
        for (int g = 1; g <= G; ++g) {
            double Egp1 = Eb1[g + 1];
            if (Egp1 > U) Egp1 = U;
            double v = (-x * erf((Watt_a * Watt_b - 2.0 * sqrt(Egp1) / (2.0 * Watt_a))) +
                        x * erf((Watt_a * Watt_b + 2.0 * sqrt(Egp1) / (2.0 * Watt_a))) -
                        2.0 * (exp(2.0 * Watt_b * sqrt(Egp1)) * exp(-(Watt_a * Watt_b * sqrt(Egp1)) / Watt_a)));
            double Eg = Eb1[g];
            if (Eg > U) Eg = U;
            v = v - (-x * erf((Watt_a * Watt_b - 2.0 * sqrt(Eg) / (2.0 * Watt_a))) +
                     x * erf((Watt_a * Watt_b + 2.0 * sqrt(Eg) / (2.0 * Watt_a))) -
                     2.0 * (exp(2.0 * Watt_b * sqrt(Eg)) * exp(-(Watt_a * Watt_b * sqrt(Eg)) / Watt_a)));
            chis[g - 1] = 0.25 * Watt_a * v / I;
        }
        break;
    }
    default:
        break;
    }
    double I = 0.0;
    for (int g = 0; g < G; ++g) I = I + chis[g];
    if (I != 1.0) {
        I = 1.0 / I;
        for (int g = 0; g < G; ++g) chis[g] = chis[g] * I;
    }
}

constexpr int CHI_MAX_SLOTS = 64;

__global__ void __launch_bounds__(128) k_chi(ChiDev c)
{
    __shared__ double s_prob[CHI_MAX_SLOTS];
    __shared__ double s_norm[CHI_MAX_SLOTS + 2];
    __shared__ double s_beta;
    const int iE = blockIdx.x;
    const int G = c.n_bins - 1;
    const double Ein = c.Ein[iE];
    const int np = c.n_prompt;
    // phase 1: one law per thread
    for (int s = threadIdx.x; s < c.n_slots; s += blockDim.x) {
        const ChiSlotDev sl = c.slots[s];
        double* out = (s < np) ? c.work + ((size_t)iE * np + s) * G : c.chi_delay + ((size_t)(s - np) * c.NE + iE) * G;
        chi_integrate(c, sl, Ein, out);
        s_prob[s] = chi_prob(c, sl, Ein);
    }
    if (threadIdx.x == blockDim.x - 1) {
        const double nd = (c.nu_d_type == 2) ? interpolate_tab1(c.nu_d_data, Ein) : 0.0;
        s_beta = nd / chi_nu_total(c, Ein);
    }
    __syncthreads();
    // phase 2: the laws combined group by group in the reference's order
    const double beta = s_beta;
    double* tot = c.chi_total + (size_t)iE * G;
    double* pr = c.chi_prompt + (size_t)iE * G;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        double p = 0.0, prob = 0.0;
        for (int s = 0; s < np; ++s) {
            prob = s_prob[s];
            p = p + prob * c.work[((size_t)iE * np + s) * G + g];
        }
        double t = p + prob * p;
        for (int s = np; s < c.n_slots; ++s) t = t + s_prob[s] * beta * c.chi_delay[((size_t)(s - np) * c.NE + iE) * G + g];
        tot[g] = t;
        pr[g] = p;
    }
    __syncthreads();
    // phase 3: normalisation, each sum taken sequentially by one thread
    const int n_sums = 2 + (c.n_slots - np);
    for (int k = threadIdx.x; k < n_sums; k += blockDim.x) {
        const double* v = (k == 0) ? tot : (k == 1) ? pr : c.chi_delay + ((size_t)(k - 2) * c.NE + iE) * G;
        double norm = 0.0;
        for (int g = 0; g < G; ++g) norm = norm + v[g];
        s_norm[k] = norm;
    }
    __syncthreads();
    for (int k = 0; k < n_sums; ++k) {
        const double norm = s_norm[k];
        if (!(norm > 0.0)) continue;
        double* v = (k == 0) ? tot : (k == 1) ? pr : c.chi_delay + ((size_t)(k - 2) * c.NE + iE) * G;
        for (int g = threadIdx.x; g < G; g += blockDim.x) v[g] = v[g] / norm;
    }
}

}  // namespace ndpp
