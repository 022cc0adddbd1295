// kernels_sab.cuh -- K6: thermal S(a,b) moments (src/sab.F90:21-454).
//
//   k_sab_el          <- integrate_sab_el         (:21-109)
//   k_sab_inel_disc   <- integrate_sab_inel_disc  (:142-245)
//   k_sab_cont_table  <- integrate_sab_inel_cont, stage 1 on the table's own E_in grid (:292-378)
//   k_sab_cont_interp <- integrate_sab_inel_cont, stage 2 (:383-407)
//   k_sab_combine     <- combine_sab_grid         (:415-454)
//
// One thread per (E_in, l).  A thread walks the outgoing energies and cosines in the reference's
// order and keeps the moment of the group it is currently scoring in a register, so every output
// element is accumulated by the same sequence of additions as in the serial Fortran loop.
// Output layout [iE][g][l].
#pragma once
#include "common.cuh"

namespace ndpp {

enum { SAB_SECONDARY_EQUAL = 0, SAB_SECONDARY_SKEWED = 1, SAB_SECONDARY_CONT = 2 };
enum { SAB_ELASTIC_DISCRETE = 3, SAB_ELASTIC_EXACT = 4 };

struct SabDev {
    double awr, kT, threshold_inelastic, threshold_elastic;
    int n_in, n_eout, n_mu, secondary_mode;
    const double* e_in; const double* sigma; const double* e_out; const double* mu;   // discrete
    const int* cont_n; const long long* cont_off; const double* cont_e; const double* cont_pdf; const double* cont_mu;
    int elastic_mode, n_el_in, n_el_mu;
    const double* el_e_in; const double* el_P; const double* el_mu;
};

// Angular basis of the output: P_l(mu) (LEGENDRE, the reference) or the indicator of the l-th of L equal-width
// cosine bins on [-1, 1] (TABULAR; the reference leaves it unimplemented, src/scatt.F90:579-588, so the
// semantics are this project's, DESIGN.md: element (b, g, E_in) = probability of group g and cosine bin b).
__device__ __forceinline__ double sab_basis(int tabular, int L, int l, double mu)
{
    if (!tabular) return calc_pn(l, mu);
    int b = (int)((mu + 1.0) * 0.5 * (double)L);
    b = b < 0 ? 0 : (b > L - 1 ? L - 1 : b);
    return (b == l) ? 1.0 : 0.0;
}

__global__ void k_sab_el(SabDev sab, const double* __restrict__ e_bins, int nbins, int L, int tabular,
                         const double* __restrict__ Ein, int NE, double* __restrict__ out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)NE * L) return;
    const int iEin = (int)(t / L), l = (int)(t % L);
    const int G = nbins - 1;
    double* col = out + (size_t)iEin * G * L;
    for (int g = 0; g < G; ++g) col[g * L + l] = 0.0;
    if (sab.threshold_elastic == 0.0) return;
    const double E = Ein[iEin];
    if (E < sab.el_e_in[0]) return;
    if (E >= sab.threshold_elastic) return;
    const int isab = binary_search(sab.el_e_in, sab.n_el_in, E);
    const double f = (E - sab.el_e_in[isab]) / (sab.el_e_in[isab + 1] - sab.el_e_in[isab]);
    if (E < e_bins[0]) return;
    if (E > e_bins[nbins - 1]) return;
    const int g = binary_search(e_bins, nbins, E);
    double sig = 0.0;
    if (sab.elastic_mode == SAB_ELASTIC_EXACT) sig = sab.el_P[isab] / E;
    else if (sab.elastic_mode == SAB_ELASTIC_DISCRETE) sig = (1.0 - f) * sab.el_P[isab] + f * sab.el_P[isab + 1];
    double v = 0.0;
    if (sab.n_el_mu == 0) {
        const double mu = 1.0 - sab.el_e_in[isab] / E;
        v = v + sab_basis(tabular, L, l, mu);
    } else if (sab.elastic_mode == SAB_ELASTIC_DISCRETE) {
        const double wgt = 1.0 / (double)sab.n_el_mu;
        for (int imu = 0; imu < sab.n_el_mu; ++imu) {
            const double mu = (1.0 - f) * sab.el_mu[(size_t)isab * sab.n_el_mu + imu] +
                              f * sab.el_mu[(size_t)(isab + 1) * sab.n_el_mu + imu];
            v = v + wgt * sab_basis(tabular, L, l, mu);
        }
    }
    col[g * L + l] = sig * v;
}

// wgt[] (n_eout) is prepared on the host exactly as :167-186
__global__ void k_sab_inel_disc(SabDev sab, const double* __restrict__ wgt, const double* __restrict__ e_bins, int nbins,
                                int L, int tabular, const double* __restrict__ Ein, int NE, double* __restrict__ out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)NE * L) return;
    const int iEin = (int)(t / L), l = (int)(t % L);
    const int G = nbins - 1, NEo = sab.n_eout, nmu = sab.n_mu;
    double* col = out + (size_t)iEin * G * L;
    for (int g = 0; g < G; ++g) col[g * L + l] = 0.0;
    const double E = Ein[iEin];
    int isab; double f;
    if (E < sab.e_in[0]) { isab = 0; f = 0.0; }
    else if (E > sab.threshold_inelastic) return;
    else if (E == sab.threshold_inelastic) { isab = sab.n_in - 2; f = 1.0; }
    else {
        isab = binary_search(sab.e_in, sab.n_in, E);
        f = (E - sab.e_in[isab]) / (sab.e_in[isab + 1] - sab.e_in[isab]);
    }
    const double sig = (1.0 - f) * sab.sigma[isab] + f * sab.sigma[isab + 1];
    int gcur = -1;
    double cur = 0.0;
    for (int iEout = 0; iEout < NEo; ++iEout) {
        const double Eout = (1.0 - f) * sab.e_out[(size_t)isab * NEo + iEout] + f * sab.e_out[(size_t)(isab + 1) * NEo + iEout];
        if (Eout < e_bins[0]) continue;
        if (Eout >= e_bins[nbins - 1]) continue;
        const int g = binary_search(e_bins, nbins, Eout);
        if (g != gcur) {
            if (gcur >= 0) col[gcur * L + l] = cur;
            gcur = g;
            cur = col[g * L + l];
        }
        const double* m0 = sab.mu + ((size_t)isab * NEo + iEout) * nmu;
        const double* m1 = sab.mu + ((size_t)(isab + 1) * NEo + iEout) * nmu;
        const double wv = wgt[iEout];
        for (int imu = 0; imu < nmu; ++imu) {
            const double mu = (1.0 - f) * m0[imu] + f * m1[imu];
            cur = cur + sab_basis(tabular, L, l, mu) * wv;
        }
    }
    if (gcur >= 0) col[gcur * L + l] = cur;
    for (int g = 0; g < G; ++g) col[g * L + l] = sig * col[g * L + l];
}

// stage 1: one thread per (table E_in row, group, l); distro[(i*G + g)*L + l]
__global__ void k_sab_cont_table(SabDev sab, const double* __restrict__ e_bins, int nbins, int L, int tabular,
                                 double* __restrict__ distro)
{
    const int G = nbins - 1, nmu = sab.n_mu;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)sab.n_in * G * L) return;
    const int l = (int)(t % L), g = (int)((t / L) % G), i = (int)(t / ((long long)L * G));
    const int NEout = sab.cont_n[i];
    const double* Eo = sab.cont_e + sab.cont_off[i];
    const double* pin = sab.cont_pdf + sab.cont_off[i];
    const double* mu_arr = sab.cont_mu + (size_t)sab.cont_off[i] * nmu;
#define SPDF(k) (((k) == NEout - 1) ? 0.0 : pin[k] * (Eo[(k) + 1] - Eo[k]))   /* :299-303 */
    double d = 0.0;
    int iE_lo, iE_hi;
    const double Eg = e_bins[g], Eg1 = e_bins[g + 1];
    bool zero = false;
    if (Eg < Eo[0]) iE_lo = 0;
    else if (Eg >= Eo[NEout - 1]) { zero = true; iE_lo = 0; }
    else {
        const int k = binary_search(Eo, NEout, Eg);
        const double f_lo = (Eg - Eo[k]) / (Eo[k + 1] - Eo[k]);
        const double mult = f_lo * SPDF(k);
        for (int imu = 0; imu < nmu; ++imu) {
            const double mu = (1.0 - f_lo) * mu_arr[(size_t)k * nmu + imu] + f_lo * mu_arr[(size_t)(k + 1) * nmu + imu];
            d = d + sab_basis(tabular, L, l, mu) * mult;
        }
        iE_lo = k + 1;
    }
    iE_hi = -1;
    if (!zero) {
        if (Eg1 < Eo[0]) zero = true;
        else if (Eg1 >= Eo[NEout - 1]) iE_hi = NEout - 2;
        else {
            const int k = binary_search(Eo, NEout, Eg1);
            const double f_hi = (Eg1 - Eo[k]) / (Eo[k + 1] - Eo[k]);
            const double mult = f_hi * SPDF(k);
            for (int imu = 0; imu < nmu; ++imu) {
                const double mu = (1.0 - f_hi) * mu_arr[(size_t)k * nmu + imu] + f_hi * mu_arr[(size_t)(k + 1) * nmu + imu];
                d = d + sab_basis(tabular, L, l, mu) * mult;
            }
            iE_hi = k - 1;
        }
    }
    if (zero) { distro[t] = 0.0; return; }
    for (int k = iE_lo; k <= iE_hi; ++k) {
        const double pk = SPDF(k);
        for (int imu = 0; imu < nmu; ++imu) d = d + sab_basis(tabular, L, l, mu_arr[(size_t)k * nmu + imu]) * pk;
    }
#undef SPDF
    distro[t] = d / (double)nmu;
}

// stage 2: one thread per output element
__global__ void k_sab_cont_interp(SabDev sab, const double* __restrict__ distro, int GL, const double* __restrict__ Ein,
                                  int NE, double* __restrict__ out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)NE * GL) return;
    const int iEin = (int)(t / GL), e = (int)(t % GL);
    const double E = Ein[iEin];
    double v;
    if (E <= sab.e_in[0]) {
        v = distro[e] * sab.sigma[0];
    } else if (E >= sab.threshold_inelastic) {
        v = 0.0;
    } else {
        const int isab = binary_search(sab.e_in, sab.n_in, E);
        const double f = (E - sab.e_in[isab]) / (sab.e_in[isab + 1] - sab.e_in[isab]);
        const double sig = (1.0 - f) * sab.sigma[isab] + f * sab.sigma[isab + 1];
        v = ((1.0 - f) * distro[(size_t)isab * GL + e] + f * distro[(size_t)(isab + 1) * GL + e]) * sig;
    }
    out[t] = v;
}

// combine: one warp per E_in.  The last column is overwritten afterwards by k_copy_last.
__global__ void k_sab_combine(const double* __restrict__ el, const double* __restrict__ inel, int G, int L, int tabular,
                              int NE, double* __restrict__ out)
{
    const int w = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= NE) return;
    const size_t o = (size_t)w * G * L;
    double norm = 0.0;
    if (tabular) {  // total probability = sum over groups and cosine bins
        for (int e = 0; e < G * L; ++e) norm = norm + (el[o + e] + inel[o + e]);
    } else {
        for (int g = 0; g < G; ++g) norm = norm + (el[o + g * L] + inel[o + g * L]);
    }
    const bool pos = norm > 0.0;
    if (pos) norm = 1.0 / norm;
    for (int e = lane; e < G * L; e += 32) out[o + e] = pos ? (el[o + e] + inel[o + e]) * norm : 0.0;
}

__global__ void k_copy_last(int GL, int NE, double* __restrict__ out)
{
    if (NE < 2) return;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < GL; e += gridDim.x * blockDim.x)
        out[(size_t)(NE - 1) * GL + e] = out[(size_t)(NE - 2) * GL + e];
}

}  // namespace ndpp
