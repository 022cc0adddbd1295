/*
 * ORACLE (test infrastructure only).  Restatement of the per-reaction integrator
 *   src/scattdata_header.F90  (convert_file4/6, integrate_file4_cm_leg, integrate_file6_cm_leg,
 *                              law9_scatter_lab_leg, integrate_file6_lab_leg, tolab, unitbase)
 *   src/search.F90:21-71      (binary_search_real)
 *   src/interpolation.F90     (interpolate_tab1)
 *   src/array_merge.F90       (merge)
 * Each function cites the lines it follows.  1-based indices throughout (A1 accessor).
 * Build: gcc -O2 -ffp-contract=off (see oracle/Makefile).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ndpp_oracle.h"

#define ZERO 0.0
#define ONE 1.0
#define TWO 2.0

/* ------------------------------------------------------------------------------------------
 * error latch: the reference's fatal_error (src/error.F90:79-154) prints and aborts.  The
 * oracle records the first message and keeps going with a clamped value so a test can report it.
 * ---------------------------------------------------------------------------------------- */
static int g_err_count = 0;
static char g_err_msg[256] = "";

void ref_fatal(const char *msg)
{
#pragma omp critical(ref_err)
    {
        if (g_err_count == 0) {
            strncpy(g_err_msg, msg, sizeof(g_err_msg) - 1);
            g_err_msg[sizeof(g_err_msg) - 1] = 0;
        }
        g_err_count++;
    }
}
int ref_error_count(void) { return g_err_count; }
const char *ref_error_message(void) { return g_err_msg; }
void ref_error_clear(void)
{
    g_err_count = 0;
    g_err_msg[0] = 0;
}

/* src/search.F90:21-71 -- returns the 1-based lower index */
int ref_binary_search(const double *array, int n, double val)
{
    int L = 1, R = n, array_index, n_iteration = 0;
    double testval;

    if (val < A1(array, L) || val > A1(array, R)) {
        ref_fatal("Value outside of array during binary search");
        return (val < A1(array, L)) ? 1 : (n > 1 ? n - 1 : 1);
    }
    while (R - L > 1) {
        if (val > A1(array, L) && val < A1(array, L + 1)) return L;
        if (val > A1(array, R - 1) && val < A1(array, R)) return R - 1;
        array_index = L + (R - L) / 2;
        testval = A1(array, array_index);
        if (val >= testval)
            L = array_index;
        else if (val < testval)
            R = array_index;
        n_iteration++;
        if (n_iteration == 64) {
            ref_fatal("Reached maximum number of iterations on binary search.");
            break;
        }
    }
    return L;
}

/* src/interpolation.F90:24-123 (array form); the object form :132-208 is the same algorithm
 * and callers flatten a Tab1 into [NR, NBT(NR), INT(NR), NP, x(NP), y(NP)]. */
double ref_interpolate_tab1(const double *data, double x)
{
    int i, j, n_regions, n_points, interp = REF_LINEAR_LINEAR;
    int loc_breakpoints, loc_interp, loc_x, loc_y;
    double r, x0, x1, y0, y1;

    n_regions = (int)A1(data, 1);
    loc_breakpoints = 1;
    loc_interp = loc_breakpoints + n_regions;
    n_points = (int)A1(data, loc_interp + n_regions + 1);
    loc_x = loc_interp + n_regions + 1;
    loc_y = loc_x + n_points;

    if (x < A1(data, loc_x + 1)) return A1(data, loc_y + 1);
    if (x > A1(data, loc_x + n_points)) return A1(data, loc_y + n_points);
    i = ref_binary_search(data + loc_x, n_points, x);

    if (n_regions == 0) {
        interp = REF_LINEAR_LINEAR;
    } else if (n_regions == 1) {
        interp = (int)A1(data, loc_interp + 1);
    } else {
        for (j = 1; j <= n_regions; ++j) {
            if (i < A1(data, loc_breakpoints + j)) {
                interp = (int)A1(data, loc_interp + j);
                break;
            }
        }
    }
    if (interp == REF_HISTOGRAM) return A1(data, loc_y + i);

    x0 = A1(data, loc_x + i);
    x1 = A1(data, loc_x + i + 1);
    y0 = A1(data, loc_y + i);
    y1 = A1(data, loc_y + i + 1);
    switch (interp) {
    case REF_LINEAR_LINEAR:
        r = (x - x0) / (x1 - x0);
        return (1 - r) * y0 + r * y1;
    case REF_LINEAR_LOG:
        r = (log(x) - log(x0)) / (log(x1) - log(x0));
        return (1 - r) * y0 + r * y1;
    case REF_LOG_LINEAR:
        r = (x - x0) / (x1 - x0);
        return exp((1 - r) * log(y0) + r * log(y1));
    case REF_LOG_LOG:
        r = (log(x) - log(x0)) / (log(x1) - log(x0));
        return exp((1 - r) * log(y0) + r * log(y1));
    default:
        ref_fatal("Unsupported interpolation scheme");
        return 0.0;
    }
}

/* src/array_merge.F90:13-107.  result must hold na+nb entries; returns the merged length. */
int ref_merge(const double *a, int na, const double *b, int nb, double *result)
{
    const double *data1, *data2;
    int ndata1, ndata2, nab, idata1, idata2, ires, no_exit = 1;

    if (A1(a, na) > A1(b, nb)) {
        data1 = b; ndata1 = nb;
        data2 = a; ndata2 = na;
    } else {
        data1 = a; ndata1 = na;
        data2 = b; ndata2 = nb;
    }
    nab = ndata1 + ndata2;
    idata1 = 1;
    idata2 = 1;
    for (ires = 1; ires <= nab; ++ires) {
        if (idata1 <= ndata1 && idata2 <= ndata2) {
            if (A1(data1, idata1) < A1(data2, idata2)) {
                A1(result, ires) = (A1(data1, idata1) == 0.0) ? REF_MIN_EIN : A1(data1, idata1);
                idata1++;
            } else if (A1(data1, idata1) == A1(data2, idata2)) {
                A1(result, ires) = A1(data1, idata1);
                idata1++;
                idata2++;
            } else {
                A1(result, ires) = (A1(data2, idata2) == 0.0) ? REF_MIN_EIN : A1(data2, idata2);
                idata2++;
            }
        } else if (idata1 <= ndata1) {
            A1(result, ires) = A1(data1, idata1);
            idata1++;
            no_exit = 0;
            break;
        } else if (idata2 <= ndata2) {
            A1(result, ires) = A1(data2, idata2);
            idata2++;
        } else {
            no_exit = 0;
            break;
        }
    }
    /* array_merge.F90:97-100: after a completed DO loop ires == nab+1 */
    if (!no_exit || ires > nab) ires = ires - 1;
    return ires;
}

/* src/scattdata_header.F90:1466-1496 */
double ref_tolab(double R, double w)
{
    double u, f;
    if (R > ONE) {
        u = (ONE + R * w) / sqrt(ONE + R * R + TWO * R * w);
    } else if (R == ONE) {
        if (w == -ONE)
            u = -ONE;
        else
            u = (ONE + R * w) / sqrt(ONE + R * R + TWO * R * w);
    } else {
        if (w < -R) {
            u = sqrt(ONE - R * R);
            f = (w - (-ONE)) / (-R - ONE);
            u = (ONE - f) * (-ONE) + f * u;
        } else {
            u = (ONE + R * w) / sqrt(ONE + R * R + TWO * R * w);
        }
    }
    return u;
}

/* src/scattdata_header.F90:669-752 (the Eouts/INTT tail :754-760 is done by the caller).
 * distro(M) must be pre-zeroed by the caller, as in scatt_convert_distro :344. */
void ref_convert_file4(int iE, const double *mu, int M, const double *ad_energy, const int *ad_type,
                       const int *ad_location, const double *data, double *distro)
{
    int lc, idata, idata_prev, imu, interp, NP;
    double r;
    (void)ad_energy;

    lc = A1(ad_location, iE);
    switch (A1(ad_type, iE)) {
    case REF_ANGLE_ISOTROPIC:
        for (imu = 1; imu <= M; ++imu) A1(distro, imu) = 0.5;
        break;
    case REF_ANGLE_32_EQUI:
        idata_prev = lc + 1;
        for (imu = 1; imu <= M; ++imu) {
            for (idata = idata_prev; idata <= lc + 1 + REF_NUM_EP; ++idata) {
                if (A1(data, idata) >= A1(mu, imu)) {
                    if (imu == 1)
                        A1(distro, imu) = REF_R_NUM_EP / (A1(data, idata + 1) - A1(data, idata));
                    else
                        A1(distro, imu) = REF_R_NUM_EP / (A1(data, idata) - A1(data, idata - 1));
                    idata_prev = idata;
                    break;
                }
            }
        }
        break;
    case REF_ANGLE_TABULAR:
        interp = (int)A1(data, lc + 1);
        NP = (int)A1(data, lc + 2);
        lc = lc + 3;
        idata_prev = lc;
        if (interp == REF_HISTOGRAM) {
            for (imu = 1; imu <= M; ++imu) {
                for (idata = idata_prev; idata <= lc + NP - 1; ++idata) {
                    if ((A1(data, idata) - A1(mu, imu)) > REF_FP_PRECISION) {
                        A1(distro, imu) = A1(data, idata - 1 + NP);
                        idata_prev = idata;
                        break;
                    } else if (fabs(A1(data, idata) - A1(mu, imu)) <= REF_FP_PRECISION) {
                        A1(distro, imu) = A1(data, idata + NP);
                        idata_prev = idata;
                        break;
                    }
                }
            }
        } else if (interp == REF_LINEAR_LINEAR) {
            for (imu = 1; imu <= M; ++imu) {
                for (idata = idata_prev; idata <= lc + NP - 1; ++idata) {
                    if ((A1(data, idata) - A1(mu, imu)) > REF_FP_PRECISION) {
                        r = (A1(mu, imu) - A1(data, idata - 1)) / (A1(data, idata) - A1(data, idata - 1));
                        A1(distro, imu) =
                            A1(data, idata + NP - 1) + r * (A1(data, idata + NP) - A1(data, idata + NP - 1));
                        idata_prev = idata;
                        break;
                    } else if (fabs(A1(data, idata) - A1(mu, imu)) <= REF_FP_PRECISION) {
                        A1(distro, imu) = A1(data, idata + NP);
                        idata_prev = idata;
                        break;
                    }
                }
            }
        }
        break;
    default:
        break; /* unknown type: distro left as is (zero) */
    }
}

/* one tabulated angular row of Law 61, src/scattdata_header.F90:843-945 */
static void ref_law61_row(const double *data, int lc, const double *mu, int M, double *col)
{
    int interp, NPang, idata, idata_prev, imu;
    double r;

    interp = (int)A1(data, lc + 1);
    NPang = (int)A1(data, lc + 2);
    lc = lc + 3;
    if (interp < REF_HISTOGRAM || interp > REF_LOG_LOG) {
        ref_fatal("Unknown interpolation type");
        return;
    }
    idata_prev = lc;
    for (imu = 1; imu <= M; ++imu) {
        for (idata = idata_prev; idata <= lc + NPang - 1; ++idata) {
            if ((A1(data, idata) - A1(mu, imu)) > REF_FP_PRECISION) {
                switch (interp) {
                case REF_HISTOGRAM:
                    A1(col, imu) = A1(data, idata + NPang - 1);
                    break;
                case REF_LINEAR_LINEAR:
                    r = (A1(mu, imu) - A1(data, idata - 1)) / (A1(data, idata) - A1(data, idata - 1));
                    A1(col, imu) =
                        A1(data, idata + NPang - 1) + r * (A1(data, idata + NPang) - A1(data, idata - 1 + NPang));
                    break;
                case REF_LINEAR_LOG:
                    r = (log(A1(mu, imu)) - log(A1(data, idata - 1))) /
                        (log(A1(data, idata)) - log(A1(data, idata - 1)));
                    A1(col, imu) =
                        A1(data, idata + NPang - 1) + r * (A1(data, idata + NPang) - A1(data, idata - 1 + NPang));
                    break;
                case REF_LOG_LINEAR:
                    r = (A1(mu, imu) - A1(data, idata - 1)) / (A1(data, idata) - A1(data, idata - 1));
                    A1(col, imu) = exp((ONE - r) * log(A1(data, idata + NPang)) + r * log(A1(data, idata + NPang - 1)));
                    break;
                case REF_LOG_LOG:
                    r = (log(A1(mu, imu)) - log(A1(data, idata - 1))) /
                        (log(A1(data, idata)) - log(A1(data, idata - 1)));
                    A1(col, imu) = exp((ONE - r) * log(A1(data, idata + NPang)) + r * log(A1(data, idata + NPang - 1)));
                    break;
                }
                idata_prev = idata;
                break;
            } else if (fabs(A1(data, idata) - A1(mu, imu)) <= REF_FP_PRECISION) {
                A1(col, imu) = A1(data, idata + NPang);
                idata_prev = idata;
                break;
            }
        }
    }
}

/* src/scattdata_header.F90:769-950.  Returns 1 (nothing touched) for an unsupported law.
 * Eouts/pdf/cdf must hold NP entries, distro M*NP (column-major, pre-zeroed). */
int ref_convert_file6(int iE, const double *mu, int M, int law, const double *data, int *INTT, int *NP_out,
                      double *Eouts, double *pdf, double *cdf, double *distro)
{
    int lcin, lc, iEout, NP, NR, NE, k;
    double KMR, KMA, KMconst;

    if (law != 4 && law != 44 && law != 61) return 1;

    NR = (int)A1(data, 1);
    if (NR > 0) {
        ref_fatal("Multiple interpolation regions not supported while attempting to sample Kalbach-Mann distribution.");
        return 2;
    }
    NE = (int)A1(data, 2 + 2 * NR);
    lc = (int)A1(data, 2 + 2 * NR + NE + iE);

    *INTT = (int)A1(data, lc + 1);
    if (*INTT > 10) *INTT = *INTT % 10;

    NP = (int)A1(data, lc + 2);
    *NP_out = NP;
    for (k = 1; k <= NP; ++k) {
        A1(Eouts, k) = A1(data, lc + 2 + k);
        A1(pdf, k) = A1(data, lc + 2 + NP + k);
        A1(cdf, k) = A1(data, lc + 2 + 2 * NP + k);
    }

    if (law == 4) {
        /* nothing else */
    } else if (law == 44) {
        lc = lc + 2;
        for (iEout = 1; iEout <= NP; ++iEout) {
            double *col = distro + (size_t)(iEout - 1) * M;
            KMR = A1(data, lc + 3 * NP + iEout);
            KMA = A1(data, lc + 4 * NP + iEout);
            KMconst = 0.5 * KMA / sinh(KMA);
            for (k = 1; k <= M; ++k)
                A1(col, k) = KMconst * (cosh(KMA * A1(mu, k)) + KMR * sinh(KMA * A1(mu, k)));
        }
    } else { /* 61 */
        lcin = lc + 2;
        for (iEout = 1; iEout <= NP; ++iEout) {
            double *col = distro + (size_t)(iEout - 1) * M;
            lc = (int)A1(data, lcin + 3 * NP + iEout);
            if (lc == 0) {
                for (k = 1; k <= M; ++k) A1(col, k) = 0.5;
                continue;
            }
            ref_law61_row(data, lc, mu, M, col);
        }
    }
    return 0;
}

/* src/scattdata_header.F90:956-1078.  distro is (order x groups) column-major; only active
 * groups are written (the caller pre-zeroes, as interp_distro/integrate_distro do). */
void ref_integrate_file4_cm_leg(const double *fw, double Ein, double awr, double Q, const double *E_bins, int nbins,
                                const double *w, int M, int order, double *distro)
{
    int g, ilo, ihi, iw, l;
    double R, wlo, whi, ulo, uhi, flo, fhi, interp, onepawr2, onepR2, inv2REin, dw;
#define D(l, g) distro[((l)-1) + (size_t)order * ((g)-1)]

    dw = A1(w, 2) - A1(w, 1);
    R = awr * sqrt((ONE + Q * (awr + ONE) / (awr * Ein)));
    onepawr2 = (ONE + awr) * (ONE + awr);
    onepR2 = ONE + R * R;
    inv2REin = 0.5 / (R * Ein);

    for (g = 1; g <= nbins - 1; ++g) {
        wlo = (A1(E_bins, g) * onepawr2 - Ein * onepR2) * inv2REin;
        if (wlo < -ONE)
            wlo = -ONE;
        else if (wlo > ONE)
            wlo = ONE;
        ilo = (int)((wlo + ONE) / dw) + 1;
        whi = (A1(E_bins, g + 1) * onepawr2 - Ein * onepR2) * inv2REin;
        if (whi < -ONE)
            whi = -ONE;
        else if (whi > ONE)
            whi = ONE;
        ihi = (int)((whi + ONE) / dw) + 1;

        if (wlo == whi) {
            if (wlo == -ONE)
                continue;
            else if (wlo == ONE)
                return;
        }

        if (ilo == M) {
            flo = A1(fw, M);
        } else {
            interp = (wlo - A1(w, ilo)) / (A1(w, ilo + 1) - A1(w, ilo));
            flo = (ONE - interp) * A1(fw, ilo) + interp * A1(fw, ilo + 1);
        }
        if (ihi == M) {
            fhi = A1(fw, M);
        } else {
            interp = (whi - A1(w, ihi)) / (A1(w, ihi + 1) - A1(w, ihi));
            fhi = (ONE - interp) * A1(fw, ihi) + interp * A1(fw, ihi + 1);
        }

        if (ilo != ihi) {
            ulo = ref_tolab(R, wlo);
            uhi = ref_tolab(R, A1(w, ilo + 1));
            for (l = 1; l <= order; ++l)
                D(l, g) = (A1(w, ilo + 1) - wlo) *
                          (flo * ref_calc_pn(l - 1, ulo) + A1(fw, ilo + 1) * ref_calc_pn(l - 1, uhi));
            for (iw = ilo + 1; iw <= ihi - 1; ++iw) {
                ulo = uhi;
                uhi = ref_tolab(R, A1(w, iw + 1));
                for (l = 1; l <= order; ++l)
                    D(l, g) = D(l, g) + (A1(w, iw + 1) - A1(w, iw)) * (A1(fw, iw) * ref_calc_pn(l - 1, ulo) +
                                                                      A1(fw, iw + 1) * ref_calc_pn(l - 1, uhi));
            }
            ulo = uhi;
            uhi = ref_tolab(R, whi);
            for (l = 1; l <= order; ++l)
                D(l, g) = D(l, g) +
                          (whi - A1(w, ihi)) * (A1(fw, ihi) * ref_calc_pn(l - 1, ulo) + fhi * ref_calc_pn(l - 1, uhi));
        } else {
            ulo = ref_tolab(R, wlo);
            uhi = ref_tolab(R, whi);
            for (l = 1; l <= order; ++l)
                D(l, g) = (whi - wlo) * (flo * ref_calc_pn(l - 1, ulo) + fhi * ref_calc_pn(l - 1, uhi));
        }
        for (l = 1; l <= order; ++l) D(l, g) = 0.5 * D(l, g);
    }
#undef D
}

/* src/scattdata_header.F90:1085-1266.  fEmu is (M x NEout) column-major. */
void ref_integrate_file6_cm_leg(const double *fEmu, const double *mu, int M, double Ein, double awr,
                                const double *Eout, int NEout, int INTT, const double *thispdf,
                                const double *E_bins, int nbins, int order, int ne_per_grp, double *distro)
{
    int g, imu_c, imu, iE, iEo, g_lo, g_hi, l;
    double Eo_cm, Eo, dEo, mu_c, mu_l_min, dmu, c, proby, f, integ, J, Eo_lo, Eo_hi, pEo, fEo, ap1inv, deltamu;
    double *fEl, *fmu, *mu_l, *pdf, *E_bnds, *seg;
#define D(l, g) distro[((l)-1) + (size_t)order * ((g)-1)]
#define F(k, j) fEmu[((k)-1) + (size_t)M * ((j)-1)]

    deltamu = A1(mu, 2) - A1(mu, 1);

    pdf = (double *)malloc(sizeof(double) * NEout);
    memcpy(pdf, thispdf, sizeof(double) * NEout);
    if (A1(Eout, NEout) == A1(Eout, NEout - 1)) A1(pdf, NEout - 1) = ZERO;

    fEl = (double *)malloc(sizeof(double) * order);
    seg = (double *)malloc(sizeof(double) * order);
    fmu = (double *)malloc(sizeof(double) * M);
    mu_l = (double *)calloc(M, sizeof(double));
    E_bnds = (double *)malloc(sizeof(double) * (nbins + 1));

    ap1inv = ONE / (awr + ONE);

    Eo_lo = A1(Eout, 1) + (Ein - TWO * (awr + ONE) * sqrt(Ein * A1(Eout, 1))) * ap1inv * ap1inv;
    Eo_lo = 1E-12; /* :1141 -- the computed bound is overwritten */
    Eo_hi = A1(Eout, NEout) + (Ein + TWO * (awr + ONE) * sqrt(Ein * A1(Eout, NEout))) * ap1inv * ap1inv;

    if (Eo_lo <= A1(E_bins, 1)) {
        g_lo = 1;
    } else if (Eo_lo >= A1(E_bins, nbins)) {
        goto done;
    } else {
        g_lo = ref_binary_search(E_bins, nbins, Eo_lo);
    }
    if (Eo_hi <= A1(E_bins, 1)) {
        goto done;
    } else if (Eo_hi >= A1(E_bins, nbins)) {
        g_hi = nbins - 1;
        A1(E_bnds, g_lo) = Eo_lo;
        for (g = g_lo + 1; g <= g_hi; ++g) A1(E_bnds, g) = A1(E_bins, g);
        A1(E_bnds, g_hi + 1) = A1(E_bins, g_hi); /* :1159 -- zero-width top group quirk */
    } else {
        g_hi = ref_binary_search(E_bins, nbins, Eo_hi);
        A1(E_bnds, g_lo) = Eo_lo;
        for (g = g_lo + 1; g <= g_hi; ++g) A1(E_bnds, g) = A1(E_bins, g);
        A1(E_bnds, g_hi + 1) = Eo_hi;
    }

    for (g = g_lo; g <= g_hi; ++g) {
        Eo = A1(E_bnds, g);
        dEo = (A1(E_bnds, g + 1) - A1(E_bnds, g)) / (double)(ne_per_grp - 1);
        Eo = Eo - dEo;
        for (iE = 1; iE <= ne_per_grp; ++iE) {
            Eo = Eo + dEo;
            for (l = 0; l < order; ++l) fEl[l] = ZERO;
            for (imu = 0; imu < M; ++imu) fmu[imu] = ZERO;
            c = ap1inv * sqrt(Ein / Eo);
            mu_l_min = (ONE + c * c - A1(Eout, NEout) / Eo) / (TWO * c);
            if (mu_l_min < -ONE) {
                mu_l_min = -ONE;
            } else if (fabs(mu_l_min - ONE) < 1E-10) {
                mu_l_min = ONE;
            } else if (mu_l_min > ONE) {
                continue;
            }
            dmu = (ONE - mu_l_min) / (double)(M - 1);
            for (imu = 1; imu <= M; ++imu) {
                A1(mu_l, imu) = mu_l_min + dmu * (double)(imu - 1);
                Eo_cm = Eo * (ONE + c * c - TWO * c * A1(mu_l, imu));
                if (Eo_cm <= ZERO) {
                    continue;
                } else if (Eo_cm <= A1(Eout, 1)) {
                    iEo = 1;
                } else if (Eo_cm >= A1(Eout, NEout)) {
                    iEo = NEout - 1;
                } else {
                    iEo = ref_binary_search(Eout, NEout, Eo_cm);
                }
                if (INTT == REF_HISTOGRAM) {
                    fEo = ZERO;
                    pEo = A1(pdf, iEo);
                } else {
                    if (A1(Eout, iEo + 1) == A1(Eout, iEo)) {
                        fEo = ZERO;
                        pEo = A1(pdf, iEo);
                    } else {
                        fEo = (Eo_cm - A1(Eout, iEo)) / (A1(Eout, iEo + 1) - A1(Eout, iEo));
                        pEo = (ONE - fEo) * A1(pdf, iEo) + fEo * A1(pdf, iEo + 1);
                    }
                }
                J = sqrt(Eo / Eo_cm);
                if (A1(mu_l, imu) == -ONE) {
                    mu_c = -ONE;
                } else if (A1(mu_l, imu) == ONE) {
                    mu_c = ONE;
                } else {
                    mu_c = (A1(mu_l, imu) - c) * J;
                    if (fabs(mu_c) > ONE) continue;
                }
                if (fabs(mu_c - ONE) < 1E-10) {
                    imu_c = M - 1;
                    f = ONE;
                } else {
                    imu_c = (int)((mu_c + ONE) / deltamu) + 1;
                    f = (mu_c - A1(mu, imu_c)) / (A1(mu, imu_c + 1) - A1(mu, imu_c));
                }
                proby = (ONE - fEo) * ((ONE - f) * F(imu_c, iEo) + f * F(imu_c + 1, iEo));
                proby = proby + fEo * ((ONE - f) * F(imu_c, iEo + 1) + f * F(imu_c + 1, iEo + 1));
                integ = proby * J * pEo;
                A1(fmu, imu) = integ;
            }
            for (imu = 1; imu <= M - 1; ++imu) {
                ref_calc_int_pn_tablelin(order, A1(mu_l, imu), A1(mu_l, imu + 1), A1(fmu, imu), A1(fmu, imu + 1), seg);
                for (l = 0; l < order; ++l) fEl[l] = fEl[l] + seg[l];
            }
            if (iE != 1 && iE != ne_per_grp) {
                for (l = 1; l <= order; ++l) D(l, g) = D(l, g) + TWO * fEl[l - 1];
            } else {
                for (l = 1; l <= order; ++l) D(l, g) = D(l, g) + fEl[l - 1];
            }
        }
        for (l = 1; l <= order; ++l) D(l, g) = D(l, g) * dEo * 0.5;
    }

    fEo = ZERO;
    for (g = g_lo; g <= g_hi; ++g) fEo = fEo + D(1, g);
    if (fEo > ZERO) fEo = ONE / fEo;
    for (g = g_lo; g <= g_hi; ++g)
        for (l = 1; l <= order; ++l) D(l, g) = D(l, g) * fEo;

done:
    free(pdf);
    free(fEl);
    free(seg);
    free(fmu);
    free(mu_l);
    free(E_bnds);
#undef D
#undef F
}

/* src/scattdata_header.F90:1274-1326 */
void ref_law9_scatter_lab_leg(const double *fmu, const double *edist_data, double Ein, const double *E_bins,
                              int nbins, const double *mu, int M, int order, double *distro)
{
    int g, NR, NE, lc, imu, l;
    double T, U, x, I, Egp1, Eg, pE_xfer;
    double *seg = (double *)malloc(sizeof(double) * order);
#define D(l, g) distro[((l)-1) + (size_t)order * ((g)-1)]

    NR = (int)A1(edist_data, 1);
    NE = (int)A1(edist_data, 2 + 2 * NR);
    T = ref_interpolate_tab1(edist_data, Ein);
    lc = 2 + 2 * NR + 2 * NE;
    U = A1(edist_data, lc + 1);
    x = (Ein - U) / T;
    I = T * T * (ONE - exp(-x) * (ONE + x));
    if (Ein - U <= ZERO) {
        free(seg);
        return;
    }
    for (g = 1; g <= nbins - 1; ++g) {
        Egp1 = A1(E_bins, g + 1);
        Eg = A1(E_bins, g);
        if (Egp1 > (Ein - U)) Egp1 = Ein - U;
        if (Eg > (Ein - U)) Eg = Ein - U;
        pE_xfer = (exp(-Egp1 / T) * (T + Egp1)) - (exp(-Eg / T) * (T + Eg));
        pE_xfer = -T * pE_xfer / I;
        for (imu = 1; imu <= M - 1; ++imu) {
            ref_calc_int_pn_tablelin(order, A1(mu, imu), A1(mu, imu + 1), A1(fmu, imu), A1(fmu, imu + 1), seg);
            for (l = 1; l <= order; ++l) D(l, g) = D(l, g) + seg[l - 1] * pE_xfer;
        }
    }
    free(seg);
#undef D
}

/* src/scattdata_header.F90:1334-1450 */
void ref_integrate_file6_lab_leg(const double *fEmu, const double *mu, int M, const double *Eout, int NEout, int INTT,
                                 const double *thispdf, const double *E_bins, int nbins, int order, double *distro)
{
    int g, imu, iE, iE_lo, iE_hi, l, groups = nbins - 1;
    double f_lo, f_hi, s;
    double *fEmu_int, *pdf, *seg;
    (void)INTT;
#define D(l, g) distro[((l)-1) + (size_t)order * ((g)-1)]
#define F(k, j) fEmu[((k)-1) + (size_t)M * ((j)-1)]
#define FI(k, g) fEmu_int[((k)-1) + (size_t)M * ((g)-1)]

    fEmu_int = (double *)calloc((size_t)M * groups, sizeof(double));
    pdf = (double *)malloc(sizeof(double) * NEout);
    seg = (double *)malloc(sizeof(double) * order);
    memcpy(pdf, thispdf, sizeof(double) * NEout);
    for (iE = 1; iE <= NEout - 1; ++iE) A1(pdf, iE) = A1(thispdf, iE) * (A1(Eout, iE + 1) - A1(Eout, iE));
    if (NEout >= 2 && A1(Eout, NEout) == A1(Eout, NEout - 1)) A1(pdf, NEout - 1) = ZERO;

    if (NEout > 1) {
        for (g = 1; g <= groups; ++g) {
            if (A1(E_bins, g) < A1(Eout, 1)) {
                iE_lo = 1;
            } else if (A1(E_bins, g) >= A1(Eout, NEout)) {
                for (l = 1; l <= order; ++l) D(l, g) = ZERO;
                continue;
            } else {
                iE_lo = ref_binary_search(Eout, NEout, A1(E_bins, g));
                f_lo = (A1(E_bins, g) - A1(Eout, iE_lo)) / (A1(Eout, iE_lo + 1) - A1(Eout, iE_lo));
                for (imu = 1; imu <= M; ++imu) FI(imu, g) = FI(imu, g) + f_lo * A1(pdf, iE_lo) * F(imu, iE_lo);
                iE_lo = iE_lo + 1;
            }
            if (A1(E_bins, g + 1) < A1(Eout, 1)) {
                for (l = 1; l <= order; ++l) D(l, g) = ZERO;
                continue;
            } else if (A1(E_bins, g + 1) >= A1(Eout, NEout)) {
                iE_hi = NEout - 1;
            } else {
                iE_hi = ref_binary_search(Eout, NEout, A1(E_bins, g + 1));
                f_hi = (A1(E_bins, g + 1) - A1(Eout, iE_hi)) / (A1(Eout, iE_hi + 1) - A1(Eout, iE_hi));
                for (imu = 1; imu <= M; ++imu) FI(imu, g) = FI(imu, g) + f_hi * A1(pdf, iE_hi) * F(imu, iE_hi);
                iE_hi = iE_hi - 1;
            }
            for (iE = iE_lo; iE <= iE_hi; ++iE)
                for (imu = 1; imu <= M; ++imu) FI(imu, g) = FI(imu, g) + A1(pdf, iE) * F(imu, iE);

            for (imu = 1; imu <= M - 1; ++imu) {
                ref_calc_int_pn_tablelin(order, A1(mu, imu), A1(mu, imu + 1), FI(imu, g), FI(imu + 1, g), seg);
                for (l = 1; l <= order; ++l) D(l, g) = D(l, g) + seg[l - 1];
            }
        }
    } else {
        for (g = 1; g <= groups; ++g) {
            if ((A1(Eout, 1) > A1(E_bins, g)) && (A1(Eout, 1) <= A1(E_bins, g + 1))) {
                for (imu = 1; imu <= M - 1; ++imu) {
                    ref_calc_int_pn_tablelin(order, A1(mu, imu), A1(mu, imu + 1), F(imu, 1), F(imu + 1, 1), seg);
                    for (l = 1; l <= order; ++l) D(l, g) = D(l, g) + seg[l - 1];
                }
            } else {
                for (l = 1; l <= order; ++l) D(l, g) = ZERO;
            }
        }
    }

    /* :1447-1448, unguarded */
    s = ZERO;
    for (g = 1; g <= groups; ++g) s = s + D(1, g);
    f_lo = ONE / s;
    for (g = 1; g <= groups; ++g)
        for (l = 1; l <= order; ++l) D(l, g) = D(l, g) * f_lo;

    free(fEmu_int);
    free(pdf);
    free(seg);
#undef D
#undef F
#undef FI
}

/* src/scattdata_header.F90:1554-1609.  ub_grid must hold n entries; returns its length. */
int ref_cast_to_unitbase(const double *Eout, int n, double *ub_grid)
{
    int ilo = 1, i, j, nt;
    double inv_dE;
    double *ub_temp = (double *)calloc(n - ilo + 1, sizeof(double));

    nt = n - ilo + 1;
    inv_dE = (A1(Eout, n) - A1(Eout, ilo));
    if ((inv_dE >= ZERO) && (inv_dE < REF_INFINITY))
        inv_dE = ONE / inv_dE;
    else
        inv_dE = ZERO;
    j = 1;
    for (i = ilo; i <= n - 1; ++i) {
        A1(ub_temp, j) = (A1(Eout, i) - A1(Eout, ilo)) * inv_dE;
        j = j + 1;
    }
    A1(ub_temp, j) = ONE;
    if (j >= 2 && A1(ub_temp, j - 1) == ONE) {
        for (i = 1; i <= nt - 1; ++i) A1(ub_grid, i) = A1(ub_temp, i);
        nt = nt - 1;
    } else {
        for (i = 1; i <= nt; ++i) A1(ub_grid, i) = A1(ub_temp, i);
    }
    free(ub_temp);
    return nt;
}

/* src/scattdata_header.F90:1616-1717.  Outputs must hold nub1+nub2 points (fEmu: M x that).
 * Returns the number of union points. */
int ref_interp_unitbase(double Ein, const double *ub1, int nub1, const double *Eout1, int n1, const double *pdf1,
                        int INTT1, const double *fEmu1, double Ei1, const double *ub2, int nub2, const double *Eout2,
                        int n2, const double *pdf2, int INTT2, const double *fEmu2, double Ei2, int M, double *Eout,
                        double *pdf, int *INTT, double *fEmu)
{
    double *ub = (double *)malloc(sizeof(double) * (nub1 + nub2));
    double p1 = 0.0, p2 = 0.0, dE1, dE2, f, r = 0.0;
    int i, j, k, nub;
    (void)INTT2; /* :1685-1697 use INTT1 for the second row as well (reference quirk) */
#define F1(k, j) fEmu1[((k)-1) + (size_t)M * ((j)-1)]
#define F2(k, j) fEmu2[((k)-1) + (size_t)M * ((j)-1)]
#define FO(k, j) fEmu[((k)-1) + (size_t)M * ((j)-1)]

    nub = ref_merge(ub1, nub1, ub2, nub2, ub);
    f = (Ein - Ei1) / (Ei2 - Ei1);
    dE1 = (A1(Eout1, n1) - A1(Eout1, 1));
    dE2 = (A1(Eout2, n2) - A1(Eout2, 1));
    for (i = 1; i <= nub; ++i) {
        j = ref_binary_search(ub1, nub1, A1(ub, i));
        if (INTT1 == REF_HISTOGRAM)
            r = ZERO;
        else if (INTT1 == REF_LINEAR_LINEAR || INTT1 == REF_LOG_LINEAR)
            r = (A1(ub, i) - A1(ub1, j)) / (A1(ub1, j + 1) - A1(ub1, j));
        else if (INTT1 == REF_LINEAR_LOG || INTT1 == REF_LOG_LOG)
            r = log(A1(ub, i) / A1(ub1, j)) / log(A1(ub1, j + 1) / A1(ub1, j));
        if (INTT1 == REF_HISTOGRAM || INTT1 == REF_LINEAR_LINEAR || INTT1 == REF_LINEAR_LOG)
            p1 = (ONE - r) * A1(pdf1, j) + r * A1(pdf1, j + 1);
        else if (INTT1 == REF_LOG_LINEAR || INTT1 == REF_LOG_LOG)
            p1 = exp((ONE - r) * log(A1(pdf1, j)) + r * log(A1(pdf1, j + 1)));
        for (k = 1; k <= M; ++k) FO(k, i) = (ONE - f) * ((ONE - r) * F1(k, j) + r * F1(k, j + 1));

        j = ref_binary_search(ub2, nub2, A1(ub, i));
        if (INTT1 == REF_HISTOGRAM)
            r = ZERO;
        else if (INTT1 == REF_LINEAR_LINEAR || INTT1 == REF_LOG_LINEAR)
            r = (A1(ub, i) - A1(ub2, j)) / (A1(ub2, j + 1) - A1(ub2, j));
        else if (INTT1 == REF_LINEAR_LOG || INTT1 == REF_LOG_LOG)
            r = log(A1(ub, i) / A1(ub2, j)) / log(A1(ub2, j + 1) / A1(ub2, j));
        if (INTT1 == REF_HISTOGRAM || INTT1 == REF_LINEAR_LINEAR || INTT1 == REF_LINEAR_LOG)
            p2 = (ONE - r) * A1(pdf2, j) + r * A1(pdf2, j + 1);
        else if (INTT1 == REF_LOG_LINEAR || INTT1 == REF_LOG_LOG)
            p2 = exp((ONE - r) * log(A1(pdf2, j)) + r * log(A1(pdf2, j + 1)));
        for (k = 1; k <= M; ++k) FO(k, i) = FO(k, i) + f * ((ONE - r) * F2(k, j) + r * F2(k, j + 1));

        A1(pdf, i) = (ONE - f) * p1 + f * p2;
        A1(Eout, i) = (ONE - f) * (A1(Eout1, 1) + dE1 * A1(ub, i)) + f * (A1(Eout2, 1) + dE2 * A1(ub, i));
    }
    *INTT = REF_LINEAR_LINEAR;
    free(ub);
    return nub;
#undef F1
#undef F2
#undef FO
}
