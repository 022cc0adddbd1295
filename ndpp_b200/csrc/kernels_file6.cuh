// kernels_file6.cuh -- K3 / K4: combined energy-angle distributions (ACE laws 4, 44, 61, 9).
//
//   k_unitbase      <- unitbase = cast_to_unitbase x2 + interp_unitbase
//                      (src/scattdata_header.F90:1521-1717) and the per-E_in scalars of
//                      scatt_interp_distro (:416-497); one thread per E_in
//   k_file6_cm      <- integrate_file6_cm_leg (:1085-1253); one block per (outgoing group, E_in)
//   k_file6_finish  <- its normalisation (:1255-1264) and the sigma * p_valid scaling (:494-497)
//   k_file6_lab     <- integrate_file6_lab_leg (:1334-1450); one block per E_in
//   k_law9          <- law9_scatter_lab_leg (:1274-1326) and its lin-lin blend (:605-638)
//
// The reference materialises the unit-base interpolated table fEmu(M, NPu) (~2-3 MB) for every
// (E_in, reaction).  Here only the union grid is stored per E_in -- NPu entries of (E_out, pdf,
// j1, r1, j2, r2) -- and fEmu(k, i) is evaluated where it is needed from the two resident table
// rows with exactly the reference's expression, so the big temporary never exists.
#pragma once
#include "common.cuh"

namespace ndpp {

struct UbDev {
    int maxU;            // stride of the per-E_in arrays
    int* n;              // [NE] union points (0 => inactive E_in)
    double* f;           // [NE] E_in interpolant between the two table rows
    InterpInfo* info;    // [NE]
    double* eout;        // [NE][maxU]
    double* pdf;         // [NE][maxU]
    int* j1;             // [NE][maxU] 0-based lower column in row iE
    double* r1;
    int* j2;             // [NE][maxU] 0-based lower column in row iE+1
    double* r2;
};

// value k of the unit-base grid of a row (cast_to_unitbase, :1592-1597)
__device__ __forceinline__ double ub_value(const double* __restrict__ Eout, int n, double inv_dE, int k)
{
    return (k == n - 1) ? 1.0 : (Eout[k] - Eout[0]) * inv_dE;
}

// binary_search (src/search.F90:21-71) over the implicit unit-base grid
__device__ __forceinline__ int ub_search(const double* __restrict__ Eout, int n, int nub, double inv_dE, double val)
{
    int L = 0, R = nub - 1;
    if (val < ub_value(Eout, n, inv_dE, L)) return 0;
    if (val > ub_value(Eout, n, inv_dE, R)) return nub > 1 ? nub - 2 : 0;
    int it = 0;
    while (R - L > 1) {
        if (val > ub_value(Eout, n, inv_dE, L) && val < ub_value(Eout, n, inv_dE, L + 1)) return L;
        if (val > ub_value(Eout, n, inv_dE, R - 1) && val < ub_value(Eout, n, inv_dE, R)) return R - 1;
        const int mid = L + (R - L) / 2;
        if (val >= ub_value(Eout, n, inv_dE, mid)) L = mid; else R = mid;
        if (++it == 64) break;
    }
    return L;
}

__device__ __forceinline__ void ub_interp(int INTT1, double u, double ua, double ub, double pa, double pb, double& r,
                                          double& p)
{
    // :1664-1677 (the reference uses INTT1 for both rows, :1685-1697)
    r = 0.0;
    if (INTT1 == LINEAR_LINEAR || INTT1 == LOG_LINEAR) r = (u - ua) / (ub - ua);
    else if (INTT1 == LINEAR_LOG || INTT1 == LOG_LOG) r = log(u / ua) / log(ub / ua);
    p = 0.0;
    if (INTT1 == HISTOGRAM || INTT1 == LINEAR_LINEAR || INTT1 == LINEAR_LOG) p = (1.0 - r) * pa + r * pb;
    else if (INTT1 == LOG_LINEAR || INTT1 == LOG_LOG) p = exp((1.0 - r) * log(pa) + r * log(pb));
}

__global__ void k_unitbase(NucDev nuc, SlotDev s, const double* __restrict__ Ein, int NE, UbDev ub, int want_ub)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NE) return;
    const double E = Ein[i];
    InterpInfo info;
    info.active = 0; info.iE = 0; info.sigS = 0.0; info.p_valid = 1.0; info.scaled = 1;
    if (E <= nuc.e_bins[nuc.n_bins - 1]) info = interp_info(nuc, s, E);
    ub.info[i] = info;
    ub.n[i] = 0;
    if (!info.active) return;
    const int iE = info.iE;
    ub.f[i] = (E - s.e_grid[iE]) / (s.e_grid[iE + 1] - s.e_grid[iE]);
    if (!want_ub) { ub.n[i] = 1; return; }

    const int oa = s.row_off[iE], ob = s.row_off[iE + 1];
    const int n1 = ob - oa, n2 = s.row_off[iE + 2] - ob;
    const double* Eo1 = s.eout + oa; const double* Eo2 = s.eout + ob;
    const double* pd1 = s.pdf + oa;  const double* pd2 = s.pdf + ob;
    const int INTT1 = s.intt[iE];
    double inv1 = Eo1[n1 - 1] - Eo1[0];
    inv1 = ((inv1 >= 0.0) && (inv1 < REF_INFINITY)) ? 1.0 / inv1 : 0.0;
    double inv2 = Eo2[n2 - 1] - Eo2[0];
    inv2 = ((inv2 >= 0.0) && (inv2 < REF_INFINITY)) ? 1.0 / inv2 : 0.0;
    int nub1 = n1, nub2 = n2;
    if (n1 >= 2 && ub_value(Eo1, n1, inv1, n1 - 2) == 1.0) nub1 = n1 - 1;
    if (n2 >= 2 && ub_value(Eo2, n2, inv2, n2 - 2) == 1.0) nub2 = n2 - 1;

    double* U = ub.eout + (size_t)i * ub.maxU;   // union grid first, converted to E_out in place
    double* P = ub.pdf + (size_t)i * ub.maxU;
    int* J1 = ub.j1 + (size_t)i * ub.maxU; int* J2 = ub.j2 + (size_t)i * ub.maxU;
    double* R1 = ub.r1 + (size_t)i * ub.maxU; double* R2 = ub.r2 + (size_t)i * ub.maxU;

    // merge (src/array_merge.F90:13-107): data1 is the array whose last entry is not larger
    const double last1 = ub_value(Eo1, n1, inv1, nub1 - 1), last2 = ub_value(Eo2, n2, inv2, nub2 - 1);
    const bool swap = last1 > last2;
    const double* A = swap ? Eo2 : Eo1; const double* B = swap ? Eo1 : Eo2;
    const int nA = swap ? n2 : n1, nB = swap ? n1 : n2, nubA = swap ? nub2 : nub1, nubB = swap ? nub1 : nub2;
    const double invA = swap ? inv2 : inv1, invB = swap ? inv1 : inv2;
    int ia = 0, ib = 0, nu = 0;
    const int nab = nubA + nubB;
    for (int k = 0; k < nab; ++k) {
        if (ia < nubA && ib < nubB) {
            const double a = ub_value(A, nA, invA, ia), b = ub_value(B, nB, invB, ib);
            if (a < b) { U[nu++] = (a == 0.0) ? MIN_EIN : a; ia++; }
            else if (a == b) { U[nu++] = a; ia++; ib++; }
            else { U[nu++] = (b == 0.0) ? MIN_EIN : b; ib++; }
        } else if (ia < nubA) {
            // "take a data1 and then stop": the stored point is dropped again by the ires
            // adjustment (:97-100), so nothing is kept
            break;
        } else if (ib < nubB) {
            U[nu++] = ub_value(B, nB, invB, ib); ib++;
        } else {
            break;
        }
    }

    const double f = ub.f[i];
    const double dE1 = Eo1[n1 - 1] - Eo1[0], dE2 = Eo2[n2 - 1] - Eo2[0];
    for (int k = 0; k < nu; ++k) {
        const double u = U[k];
        double r, p1, p2;
        int j = ub_search(Eo1, n1, nub1, inv1, u);
        ub_interp(INTT1, u, ub_value(Eo1, n1, inv1, j), ub_value(Eo1, n1, inv1, j + 1), pd1[j], pd1[j + 1], r, p1);
        J1[k] = j; R1[k] = r;
        j = ub_search(Eo2, n2, nub2, inv2, u);
        ub_interp(INTT1, u, ub_value(Eo2, n2, inv2, j), ub_value(Eo2, n2, inv2, j + 1), pd2[j], pd2[j + 1], r, p2);
        J2[k] = j; R2[k] = r;
        P[k] = (1.0 - f) * p1 + f * p2;
        U[k] = (1.0 - f) * (Eo1[0] + dE1 * u) + f * (Eo2[0] + dE2 * u);
    }
    ub.n[i] = nu;
}

// View of the unit-base interpolated table of one E_in, evaluated on demand (:1679-1702).
struct FEmu {
    const double* F1; const double* F2;  // rows iE and iE+1 of the slot's tables
    const int* j1; const int* j2; const double* r1; const double* r2;
    double f; int M;
    __device__ __forceinline__ double operator()(int k, int i) const
    {
        const double* c1 = F1 + (size_t)j1[i] * M + k;
        const double* c2 = F2 + (size_t)j2[i] * M + k;
        double v = (1.0 - f) * ((1.0 - r1[i]) * c1[0] + r1[i] * c1[M]);
        v = v + f * ((1.0 - r2[i]) * c2[0] + r2[i] * c2[M]);
        return v;
    }
};

// integrate_file6_cm_leg: block (g, iEin).  raw[(iEin*G + g)*L + l] receives the un-normalised
// group moments.  Warps take the NE_PER_GRP outgoing energies of the group round-robin; lanes run
// over the M lab cosines in windows of 32 points / 31 segments.
// Dynamic shared memory: maxU*(4 doubles + 2 ints) + ne_per_grp*L doubles.
#ifndef NDPP_F6_MINBLOCKS
#define NDPP_F6_MINBLOCKS 4
#endif
template <int LT>
__global__ void __launch_bounds__(128, NDPP_F6_MINBLOCKS) k_file6_cm(NucDev nuc, SlotDev s, const double* __restrict__ Ein, UbDev ub, double* __restrict__ raw)
{
    extern __shared__ double sm[];
    const int g = blockIdx.y, iEin = blockIdx.x;  // E_in on x: up to 2^31-1 columns
    const int NPu = ub.n[iEin];
    if (NPu == 0) return;
    constexpr int L = LT;
    const int M = nuc.M, K = nuc.ne_per_grp, nbins = nuc.n_bins;
    const double E = Ein[iEin];
    const double awr = nuc.awr;

    double* eo = sm;
    double* pd = eo + ub.maxU;
    double* r1 = pd + ub.maxU;
    double* r2 = r1 + ub.maxU;
    double* items = r2 + ub.maxU;               // [K][L]
    int* j1 = (int*)(items + (size_t)K * L);
    int* j2 = j1 + ub.maxU;

    // group range (:1138-1166); Eo_lo is overwritten by 1e-12 in the reference (:1141)
    const double* Eg = ub.eout + (size_t)iEin * ub.maxU;
    const double Eout_last = Eg[NPu - 1];
    const double ap1inv = 1.0 / (awr + 1.0);
    const double Eo_lo = 1E-12;
    const double Eo_hi = Eout_last + (E + 2.0 * (awr + 1.0) * sqrt(E * Eout_last)) * ap1inv * ap1inv;
    int g_lo, g_hi;  // 0-based
    double top;      // E_bnds(g_hi + 1)
    if (Eo_lo <= nuc.e_bins[0]) g_lo = 0;
    else if (Eo_lo >= nuc.e_bins[nbins - 1]) return;
    else g_lo = binary_search(nuc.e_bins, nbins, Eo_lo);
    if (Eo_hi <= nuc.e_bins[0]) return;
    else if (Eo_hi >= nuc.e_bins[nbins - 1]) { g_hi = nbins - 2; top = nuc.e_bins[g_hi]; }  // :1159 quirk
    else { g_hi = binary_search(nuc.e_bins, nbins, Eo_hi); top = Eo_hi; }
    if (g < g_lo || g > g_hi) return;
    const double Eb_lo = (g == g_lo) ? Eo_lo : nuc.e_bins[g];
    const double Eb_hi = (g == g_hi) ? top : nuc.e_bins[g + 1];

    for (int k = threadIdx.x; k < NPu; k += blockDim.x) {
        eo[k] = Eg[k];
        pd[k] = ub.pdf[(size_t)iEin * ub.maxU + k];
        r1[k] = ub.r1[(size_t)iEin * ub.maxU + k];
        r2[k] = ub.r2[(size_t)iEin * ub.maxU + k];
        j1[k] = ub.j1[(size_t)iEin * ub.maxU + k];
        j2[k] = ub.j2[(size_t)iEin * ub.maxU + k];
    }
    __syncthreads();
    if (threadIdx.x == 0 && NPu >= 2 && eo[NPu - 1] == eo[NPu - 2]) pd[NPu - 2] = 0.0;  // :1127-1130
    __syncthreads();

    const int iE = ub.info[iEin].iE;
    FEmu F;
    F.F1 = s.tab + (size_t)s.row_off[iE] * M; F.F2 = s.tab + (size_t)s.row_off[iE + 1] * M;
    F.j1 = j1; F.j2 = j2; F.r1 = r1; F.r2 = r2; F.f = ub.f[iEin]; F.M = M;

    const double* __restrict__ mu = nuc.mu;
    const double deltamu = mu[1] - mu[0];
    const double dEo = (Eb_hi - Eb_lo) / (double)(K - 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;

    for (int it = warp; it < K; it += nw) {
        // Eo accumulated as the reference accumulates it (:1169-1173)
        double Eo = Eb_lo - dEo;
        for (int q = 0; q <= it; ++q) Eo = Eo + dEo;
        double fEl[NDPP_MAX_L];
#pragma unroll
        for (int l = 0; l < NDPP_MAX_L; ++l) fEl[l] = 0.0;
        const double c = ap1inv * sqrt(E / Eo);
        double mu_l_min = (1.0 + c * c - Eout_last / Eo) / (2.0 * c);
        bool skip = false;
        if (mu_l_min < -1.0) mu_l_min = -1.0;
        else if (fabs(mu_l_min - 1.0) < 1E-10) mu_l_min = 1.0;
        else if (mu_l_min > 1.0) skip = true;
        if (!skip) {
            const double dmu = (1.0 - mu_l_min) / (double)(M - 1);
            for (int base = 0; base < M - 1; base += 31) {
                const int p = base + lane;
                double x = 0.0, fv = 0.0;
                if (p < M) {
                    x = mu_l_min + dmu * (double)p;
                    const double Eo_cm = Eo * (1.0 + c * c - 2.0 * c * x);
                    if (Eo_cm > 0.0) {
                        int iEo;
                        if (Eo_cm <= eo[0]) iEo = 0;
                        else if (Eo_cm >= eo[NPu - 1]) iEo = NPu - 2;
                        else iEo = binary_search_nonneg(eo, NPu, Eo_cm);
                        double fEo, pEo;
                        // the reference's INTT after unit-base interpolation is always lin-lin (:1716)
                        if (eo[iEo + 1] == eo[iEo]) { fEo = 0.0; pEo = pd[iEo]; }
                        else {
                            fEo = (Eo_cm - eo[iEo]) / (eo[iEo + 1] - eo[iEo]);
                            pEo = (1.0 - fEo) * pd[iEo] + fEo * pd[iEo + 1];
                        }
                        const double J = sqrt(Eo / Eo_cm);
                        double mu_c;
                        bool ok = true;
                        if (x == -1.0) mu_c = -1.0;
                        else if (x == 1.0) mu_c = 1.0;
                        else { mu_c = (x - c) * J; if (fabs(mu_c) > 1.0) ok = false; }
                        if (ok) {
                            int k0; double ff;
                            if (fabs(mu_c - 1.0) < 1E-10) { k0 = M - 2; ff = 1.0; }
                            else {
                                k0 = (int)((mu_c + 1.0) / deltamu);
                                ff = (mu_c - mu[k0]) / (mu[k0 + 1] - mu[k0]);
                            }
                            double proby = (1.0 - fEo) * ((1.0 - ff) * F(k0, iEo) + ff * F(k0 + 1, iEo));
                            proby = proby + fEo * ((1.0 - ff) * F(k0, iEo + 1) + ff * F(k0 + 1, iEo + 1));
                            fv = proby * J * pEo;
                        }
                    }
                }
                const double fnext = __shfl_down_sync(0xffffffffu, fv, 1);
                // a segment whose two end values are zero adds exact zeros: skipped
                if (lane < 31 && p + 1 < M && (fv != 0.0 || fnext != 0.0)) {
                    const double xh = mu_l_min + dmu * (double)(p + 1);
                    Powers A, B;
                    make_powers(x, A);
                    make_powers(xh, B);
                    add_int_pn_tablelin<LT>(L, x, xh, fv, fnext, A, B, fEl);
                }
            }
        }
#pragma unroll
        for (int l = 0; l < NDPP_MAX_L; ++l) {
            if (l < L) {
                const double v = warp_sum(fEl[l]);
                if (lane == 0) items[it * L + l] = v;
            }
        }
    }
    __syncthreads();
    // trapezoid over the K outgoing energies, in the reference's order (:1246-1252)
    if (threadIdx.x < L) {
        const int l = threadIdx.x;
        double d = 0.0;
        for (int it = 0; it < K; ++it) {
            const double v = items[it * L + l];
            d = (it != 0 && it != K - 1) ? d + 2.0 * v : d + v;
        }
        raw[((size_t)iEin * nuc.G + g) * L + l] = d * dEo * 0.5;
    }
}

// Normalisation over the groups (:1255-1264; raw is zero outside g_lo..g_hi, so summing over all
// groups in order gives the same value) followed by distro * sigS * p_valid (:494-497).
// One warp per E_in; out may alias raw.
// raw and out may alias (the callers finish a slab in place): no __restrict__ on either.
__global__ void k_file6_finish(NucDev nuc, UbDev ub, int NE, const double* raw, double* out, int guard)
{
    const int w = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= NE) return;
    const int G = nuc.G, L = nuc.L;
    const double* r = raw + (size_t)w * G * L;
    double* o = out + (size_t)w * G * L;
    const InterpInfo info = ub.info[w];
    if (!info.active) {
        for (int e = lane; e < G * L; e += 32) o[e] = 0.0;
        return;
    }
    double fEo = 0.0;
    for (int g = 0; g < G; ++g) fEo = fEo + r[g * L];
    if (guard) { if (fEo > 0.0) fEo = 1.0 / fEo; }   // file6_cm: guarded (:1261)
    else fEo = 1.0 / fEo;                             // file6_lab: unguarded (:1447)
    __syncwarp();
    for (int e = lane; e < G * L; e += 32) o[e] = apply_scale(info, r[e] * fEo);
}

// integrate_file6_lab_leg: one block per E_in, warps over groups.
// Dynamic shared memory: maxU*(4 doubles + 2 ints) + maxU doubles (group-integrated pdf).
__global__ void k_file6_lab(NucDev nuc, SlotDev s, const double* __restrict__ Ein, UbDev ub, double* __restrict__ raw)
{
    extern __shared__ double sm[];
    const int iEin = blockIdx.x;
    const int NPu = ub.n[iEin];
    const int M = nuc.M, L = nuc.L, G = nuc.G;
    double* out = raw + (size_t)iEin * G * L;
    if (NPu == 0) return;

    double* eo = sm;
    double* pd = eo + ub.maxU;
    double* r1 = pd + ub.maxU;
    double* r2 = r1 + ub.maxU;
    double* pw = r2 + ub.maxU;                  // pdf(i) * (Eout(i+1) - Eout(i)), :1361-1367
    int* j1 = (int*)(pw + ub.maxU);
    int* j2 = j1 + ub.maxU;
    const size_t o = (size_t)iEin * ub.maxU;
    for (int k = threadIdx.x; k < NPu; k += blockDim.x) {
        eo[k] = ub.eout[o + k]; pd[k] = ub.pdf[o + k];
        r1[k] = ub.r1[o + k]; r2[k] = ub.r2[o + k]; j1[k] = ub.j1[o + k]; j2[k] = ub.j2[o + k];
    }
    __syncthreads();
    for (int k = threadIdx.x; k < NPu; k += blockDim.x) pw[k] = (k < NPu - 1) ? pd[k] * (eo[k + 1] - eo[k]) : pd[k];
    __syncthreads();
    if (threadIdx.x == 0 && NPu >= 2 && eo[NPu - 1] == eo[NPu - 2]) pw[NPu - 2] = 0.0;
    __syncthreads();

    const int iE = ub.info[iEin].iE;
    FEmu F;
    F.F1 = s.tab + (size_t)s.row_off[iE] * M; F.F2 = s.tab + (size_t)s.row_off[iE + 1] * M;
    F.j1 = j1; F.j2 = j2; F.r1 = r1; F.r2 = r2; F.f = ub.f[iEin]; F.M = M;
    const double* __restrict__ mu = nuc.mu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;

    for (int g = warp; g < G; g += nw) {
        const double Elo = nuc.e_bins[g], Ehi = nuc.e_bins[g + 1];
        // term list of fEmu_int(:, g) in the reference's order (:1376-1418)
        int lo_col = -1, hi_col = -1, iE_lo = 0, iE_hi = -1;
        double lo_w = 0.0, hi_w = 0.0;
        bool zero = false, single = false;
        if (NPu > 1) {
            if (Elo < eo[0]) iE_lo = 0;
            else if (Elo >= eo[NPu - 1]) zero = true;
            else {
                const int k = binary_search(eo, NPu, Elo);
                const double f_lo = (Elo - eo[k]) / (eo[k + 1] - eo[k]);
                lo_col = k; lo_w = f_lo * pw[k];
                iE_lo = k + 1;
            }
            if (!zero) {
                if (Ehi < eo[0]) zero = true;
                else if (Ehi >= eo[NPu - 1]) iE_hi = NPu - 2;
                else {
                    const int k = binary_search(eo, NPu, Ehi);
                    const double f_hi = (Ehi - eo[k]) / (eo[k + 1] - eo[k]);
                    hi_col = k; hi_w = f_hi * pw[k];
                    iE_hi = k - 1;
                }
            }
        } else {
            single = true;
            zero = !((eo[0] > Elo) && (eo[0] <= Ehi));  // :1433
        }
        double acc[NDPP_MAX_L];
#pragma unroll
        for (int l = 0; l < NDPP_MAX_L; ++l) acc[l] = 0.0;
        if (!zero) {
            for (int base = 0; base < M - 1; base += 31) {
                const int p = base + lane;
                double fv = 0.0;
                if (p < M) {
                    if (single) fv = F(p, 0);
                    else {
                        if (lo_col >= 0) fv = fv + lo_w * F(p, lo_col);
                        if (hi_col >= 0) fv = fv + hi_w * F(p, hi_col);
                        for (int k = iE_lo; k <= iE_hi; ++k) fv = fv + pw[k] * F(p, k);
                    }
                }
                const double fnext = __shfl_down_sync(0xffffffffu, fv, 1);
                if (lane < 31 && p + 1 < M) {
                    Powers A, B;
                    make_powers(mu[p], A);
                    make_powers(mu[p + 1], B);
                    add_int_pn_tablelin(L, mu[p], mu[p + 1], fv, fnext, A, B, acc);
                }
            }
        }
#pragma unroll
        for (int l = 0; l < NDPP_MAX_L; ++l) {
            if (l < L) {
                const double v = warp_sum(acc[l]);
                if (lane == 0) out[g * L + l] = v;
            }
        }
    }
}

// law9_scatter_lab_leg for both table rows + the lin-lin blend; one block per E_in.
// out[iEin][g][l] receives the blended, scaled result.
__global__ void k_law9(NucDev nuc, SlotDev s, const double* __restrict__ Ein, UbDev ub, double* __restrict__ out)
{
    __shared__ double mom[2][NDPP_MAX_L];
    __shared__ double part[2][NDPP_MAX_L][32];
    const int iEin = blockIdx.x;
    const int M = nuc.M, L = nuc.L, G = nuc.G;
    double* o = out + (size_t)iEin * G * L;
    const InterpInfo info = ub.info[iEin];
    if (!info.active) {
        for (int e = threadIdx.x; e < G * L; e += blockDim.x) o[e] = 0.0;
        return;
    }
    const double E = Ein[iEin];
    const int iE = info.iE;
    const double* __restrict__ mu = nuc.mu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    // Legendre moments of the two rows over the whole mu grid (:1319-1323)
    for (int row = 0; row < 2; ++row) {
        const double* fm = s.tab + (size_t)s.row_off[iE + row] * M;
        double acc[NDPP_MAX_L];
#pragma unroll
        for (int l = 0; l < NDPP_MAX_L; ++l) acc[l] = 0.0;
        for (int p = threadIdx.x; p < M - 1; p += blockDim.x) {
            Powers A, B;
            make_powers(mu[p], A);
            make_powers(mu[p + 1], B);
            add_int_pn_tablelin(L, mu[p], mu[p + 1], fm[p], fm[p + 1], A, B, acc);
        }
#pragma unroll
        for (int l = 0; l < NDPP_MAX_L; ++l) {
            if (l < L) {
                const double v = warp_sum(acc[l]);
                if (lane == 0) part[row][l][warp] = v;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 2 * L) {
        const int row = threadIdx.x / L, l = threadIdx.x % L;
        double v = 0.0;
        for (int q = 0; q < nw; ++q) v = v + part[row][l][q];
        mom[row][l] = v;
    }
    __syncthreads();
    // group probabilities of the evaporation spectrum (:1290-1312)
    const double* data = s.ed_data;
    const int NR = (int)data[0];
    const int NEd = (int)data[1 + 2 * NR];
    const double T = interpolate_tab1(data, E);
    const double U = data[2 + 2 * NR + 2 * NEd];
    const double x = (E - U) / T;
    const double I = T * T * (1.0 - exp(-x) * (1.0 + x));
    const double f = ub.f[iEin];
    for (int e = threadIdx.x; e < G * L; e += blockDim.x) {
        const int g = e / L, l = e % L;
        double v = 0.0;
        if (!(E - U <= 0.0)) {
            double Egp1 = nuc.e_bins[g + 1], Eg = nuc.e_bins[g];
            if (Egp1 > (E - U)) Egp1 = E - U;
            if (Eg > (E - U)) Eg = E - U;
            double pE = (exp(-Egp1 / T) * (T + Egp1)) - (exp(-Eg / T) * (T + Eg));
            pE = -T * pE / I;
            v = (1.0 - f) * (mom[0][l] * pE) + f * (mom[1][l] * pE);
        }
        o[e] = apply_scale(info, v);
    }
}

}  // namespace ndpp
