/*
 * ORACLE (test infrastructure only -- never linked into or called by the product path).
 *
 * CPU restatement of the reference's S(a,b) thermal moments:
 *   integrate_sab_el         src/sab.F90:21-109
 *   integrate_sab_inel_disc  src/sab.F90:142-245
 *   integrate_sab_inel_cont  src/sab.F90:253-408
 *   combine_sab_grid         src/sab.F90:415-454
 *   calc_scattsab            src/scatt.F90:543-596  (Legendre branch)
 * Serial semantics (the reference's OpenMP build races on `sig`).
 * Parity status: no reference test exists for this module => "parity unpinned"; tests hold every
 * routine against an independent numpy evaluation of the Fortran text (tests/test_oracle_golden.py).
 *
 * Flattening of type(SAlphaBeta) (src/ace_header.F90:201-235), Fortran column-major kept:
 *   inelastic_e_out(NEo,NEi)      -> e_out[(isab-1)*NEo + (iEout-1)]
 *   inelastic_mu(n_mu,NEo,NEi)    -> mu[((isab-1)*NEo + (iEout-1))*n_mu + (imu-1)]
 *   inelastic_data(i) (continuous)-> cont_n_e_out[i], and cont_e_out / cont_pdf / cont_mu
 *                                    concatenated row after row (mu(n_mu, NEo_i) column-major)
 *   elastic_mu(n_mu,NEe)          -> elastic_mu[(isab-1)*n_mu + (imu-1)]
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "ndpp_oracle.h"

#define ZERO 0.0
#define ONE 1.0

void ref_fatal(const char *msg);

typedef struct {
    double awr, kT, threshold_inelastic, threshold_elastic;
    int n_inelastic_e_in, n_inelastic_e_out, n_inelastic_mu, secondary_mode;
    double *inelastic_e_in, *inelastic_sigma, *inelastic_e_out, *inelastic_mu;
    int *cont_n_e_out;
    long *cont_off; /* offset of row i in cont_e_out/cont_pdf (points) */
    double *cont_e_out, *cont_pdf, *cont_mu;
    int elastic_mode, n_elastic_e_in, n_elastic_mu;
    double *elastic_e_in, *elastic_P, *elastic_mu;
} ref_sab;

static double *dupd(const double *src, size_t n)
{
    double *p;
    if (!src || n == 0) return NULL;
    p = (double *)malloc(n * sizeof(double));
    memcpy(p, src, n * sizeof(double));
    return p;
}

void *ref_sab_create(double awr, double kT, double threshold_inelastic, double threshold_elastic, int n_inelastic_e_in,
                     int n_inelastic_e_out, int n_inelastic_mu, int secondary_mode, const double *inelastic_e_in,
                     const double *inelastic_sigma, const double *inelastic_e_out, const double *inelastic_mu,
                     const int *cont_n_e_out, const double *cont_e_out, const double *cont_pdf, const double *cont_mu,
                     int elastic_mode, int n_elastic_e_in, int n_elastic_mu, const double *elastic_e_in,
                     const double *elastic_P, const double *elastic_mu)
{
    ref_sab *s = (ref_sab *)calloc(1, sizeof(ref_sab));
    int i;
    long tot = 0;
    s->awr = awr;
    s->kT = kT;
    s->threshold_inelastic = threshold_inelastic;
    s->threshold_elastic = threshold_elastic;
    s->n_inelastic_e_in = n_inelastic_e_in;
    s->n_inelastic_e_out = n_inelastic_e_out;
    s->n_inelastic_mu = n_inelastic_mu;
    s->secondary_mode = secondary_mode;
    s->inelastic_e_in = dupd(inelastic_e_in, n_inelastic_e_in);
    s->inelastic_sigma = dupd(inelastic_sigma, n_inelastic_e_in);
    if (secondary_mode == REF_SAB_SECONDARY_CONT) {
        s->cont_n_e_out = (int *)malloc(sizeof(int) * n_inelastic_e_in);
        s->cont_off = (long *)malloc(sizeof(long) * (n_inelastic_e_in + 1));
        for (i = 0; i < n_inelastic_e_in; ++i) {
            s->cont_n_e_out[i] = cont_n_e_out[i];
            s->cont_off[i] = tot;
            tot += cont_n_e_out[i];
        }
        s->cont_off[n_inelastic_e_in] = tot;
        s->cont_e_out = dupd(cont_e_out, tot);
        s->cont_pdf = dupd(cont_pdf, tot);
        s->cont_mu = dupd(cont_mu, (size_t)tot * n_inelastic_mu);
    } else {
        s->inelastic_e_out = dupd(inelastic_e_out, (size_t)n_inelastic_e_out * n_inelastic_e_in);
        s->inelastic_mu = dupd(inelastic_mu, (size_t)n_inelastic_mu * n_inelastic_e_out * n_inelastic_e_in);
    }
    s->elastic_mode = elastic_mode;
    s->n_elastic_e_in = n_elastic_e_in;
    s->n_elastic_mu = n_elastic_mu;
    s->elastic_e_in = dupd(elastic_e_in, n_elastic_e_in);
    s->elastic_P = dupd(elastic_P, n_elastic_e_in);
    s->elastic_mu = dupd(elastic_mu, (size_t)n_elastic_mu * n_elastic_e_in);
    return s;
}

void ref_sab_free(void *h)
{
    ref_sab *s = (ref_sab *)h;
    if (!s) return;
    free(s->inelastic_e_in); free(s->inelastic_sigma); free(s->inelastic_e_out); free(s->inelastic_mu);
    free(s->cont_n_e_out); free(s->cont_off); free(s->cont_e_out); free(s->cont_pdf); free(s->cont_mu);
    free(s->elastic_e_in); free(s->elastic_P); free(s->elastic_mu);
    free(s);
}

/* Angular basis of the output.  LEGENDRE: P_l(mu), l = 0..order (the reference).  TABULAR: the reference has
 * no implementation (calc_scattsab: "TODO", src/scatt.F90:579-588; scattdata_header.F90:658-660), so the
 * semantics are this project's: `order` equal-width cosine bins on [-1, 1]; a discrete cosine
 * deposits its weight in the bin that contains it (mu = 1 in the last bin), so that element (b, g, E_in) is the
 * probability of scattering into group g and cosine bin b.  Parity for this mode is unpinned. */
static int g_sab_tabular = 0;
static int sab_orders(int order) { return g_sab_tabular ? order : order + 1; }
static double sab_basis(int L, int l, double mu)
{
    int b;
    if (!g_sab_tabular) return ref_calc_pn(l, mu);
    b = (int)((mu + ONE) * 0.5 * (double)L);
    if (b < 0) b = 0;
    if (b > L - 1) b = L - 1;
    return (b == l) ? ONE : ZERO;
}

#define SI(l, g, iE) sab_int[((l)-1) + (size_t)L * (((g)-1) + (size_t)groups * ((iE)-1))]

/* src/sab.F90:21-109 */
static void integrate_sab_el(const ref_sab *sab, const double *ein_grid, int NE, const double *e_bins, int nbins,
                             int order, double *sab_int)
{
    int L = sab_orders(order), groups = nbins - 1, iEin, g, imu, l, isab;
    double sig = ZERO, mu, Ein, f, wgt = ZERO;

    memset(sab_int, 0, sizeof(double) * (size_t)L * groups * NE);
    if (sab->threshold_elastic == ZERO) return;
    if (sab->elastic_mode == REF_SAB_ELASTIC_DISCRETE) wgt = ONE / (double)sab->n_elastic_mu;

    for (iEin = 1; iEin <= NE; ++iEin) {
        Ein = A1(ein_grid, iEin);
        if (Ein < A1(sab->elastic_e_in, 1)) {
            continue;
        } else if (Ein >= sab->threshold_elastic) {
            continue;
        } else {
            isab = ref_binary_search(sab->elastic_e_in, sab->n_elastic_e_in, Ein);
            f = (Ein - A1(sab->elastic_e_in, isab)) / (A1(sab->elastic_e_in, isab + 1) - A1(sab->elastic_e_in, isab));
        }
        if (Ein < A1(e_bins, 1)) {
            continue;
        } else if (Ein > A1(e_bins, nbins)) {
            continue;
        } else {
            g = ref_binary_search(e_bins, nbins, Ein);
        }
        if (sab->elastic_mode == REF_SAB_ELASTIC_EXACT)
            sig = A1(sab->elastic_P, isab) / Ein;
        else if (sab->elastic_mode == REF_SAB_ELASTIC_DISCRETE)
            sig = (ONE - f) * A1(sab->elastic_P, isab) + f * A1(sab->elastic_P, isab + 1);

        if (sab->n_elastic_mu == 0) {
            mu = ONE - A1(sab->elastic_e_in, isab) / Ein;
            for (l = 1; l <= L; ++l) SI(l, g, iEin) = SI(l, g, iEin) + sab_basis(L, l - 1, mu);
        } else if (sab->elastic_mode == REF_SAB_ELASTIC_DISCRETE) {
            for (imu = 1; imu <= sab->n_elastic_mu; ++imu) {
                mu = (ONE - f) * sab->elastic_mu[(size_t)(isab - 1) * sab->n_elastic_mu + (imu - 1)] +
                     f * sab->elastic_mu[(size_t)isab * sab->n_elastic_mu + (imu - 1)];
                for (l = 1; l <= L; ++l) SI(l, g, iEin) = SI(l, g, iEin) + wgt * sab_basis(L, l - 1, mu);
            }
        }
        for (g = 1; g <= groups; ++g)
            for (l = 1; l <= L; ++l) SI(l, g, iEin) = sig * SI(l, g, iEin);
    }
}

/* src/sab.F90:142-245 */
static void integrate_sab_inel_disc(const ref_sab *sab, const double *ein_grid, int NE, const double *e_bins,
                                    int nbins, int order, double *sab_int)
{
    int L = sab_orders(order), groups = nbins - 1, iEin, iEout, g, imu, l, isab, NEo = sab->n_inelastic_e_out,
        nmu = sab->n_inelastic_mu;
    double sig, mu, Ein, Eout, f, s;
    double *wgt = (double *)malloc(sizeof(double) * (NEo > 0 ? NEo : 1));

    memset(sab_int, 0, sizeof(double) * (size_t)L * groups * NE);
    if (sab->secondary_mode == REF_SAB_SECONDARY_EQUAL) {
        for (iEout = 1; iEout <= NEo; ++iEout) A1(wgt, iEout) = ONE / ((double)NEo * (double)nmu);
    } else {
        if (NEo > 4) {
            A1(wgt, 1) = 0.1;
            A1(wgt, 2) = 0.4;
            for (iEout = 3; iEout <= NEo - 2; ++iEout) A1(wgt, iEout) = ONE;
            A1(wgt, NEo - 1) = 0.4;
            A1(wgt, NEo) = 0.1;
            s = ZERO;
            for (iEout = 1; iEout <= NEo; ++iEout) s = s + A1(wgt, iEout);
            for (iEout = 1; iEout <= NEo; ++iEout) A1(wgt, iEout) = A1(wgt, iEout) / (s * (double)nmu);
        } else {
            ref_fatal("Number of Inelastic Outgoing Energies Less Than 4, but Skewed Weighting Requested by Data!");
            free(wgt);
            return;
        }
    }

    for (iEin = 1; iEin <= NE; ++iEin) {
        Ein = A1(ein_grid, iEin);
        if (Ein < A1(sab->inelastic_e_in, 1)) {
            isab = 1;
            f = ZERO;
        } else if (Ein > sab->threshold_inelastic) {
            continue;
        } else if (Ein == sab->threshold_inelastic) {
            isab = sab->n_inelastic_e_in - 1;
            f = ONE;
        } else {
            isab = ref_binary_search(sab->inelastic_e_in, sab->n_inelastic_e_in, Ein);
            f = (Ein - A1(sab->inelastic_e_in, isab)) /
                (A1(sab->inelastic_e_in, isab + 1) - A1(sab->inelastic_e_in, isab));
        }
        sig = (ONE - f) * A1(sab->inelastic_sigma, isab) + f * A1(sab->inelastic_sigma, isab + 1);

        for (iEout = 1; iEout <= NEo; ++iEout) {
            Eout = (ONE - f) * sab->inelastic_e_out[(size_t)(isab - 1) * NEo + (iEout - 1)] +
                   f * sab->inelastic_e_out[(size_t)isab * NEo + (iEout - 1)];
            if (Eout < A1(e_bins, 1)) {
                continue;
            } else if (Eout >= A1(e_bins, nbins)) {
                continue;
            } else {
                g = ref_binary_search(e_bins, nbins, Eout);
            }
            for (imu = 1; imu <= nmu; ++imu) {
                mu = (ONE - f) * sab->inelastic_mu[((size_t)(isab - 1) * NEo + (iEout - 1)) * nmu + (imu - 1)] +
                     f * sab->inelastic_mu[((size_t)isab * NEo + (iEout - 1)) * nmu + (imu - 1)];
                for (l = 1; l <= L; ++l) SI(l, g, iEin) = SI(l, g, iEin) + sab_basis(L, l - 1, mu) * A1(wgt, iEout);
            }
        }
        for (g = 1; g <= groups; ++g)
            for (l = 1; l <= L; ++l) SI(l, g, iEin) = sig * SI(l, g, iEin);
    }
    free(wgt);
}

/* src/sab.F90:253-408 */
static void integrate_sab_inel_cont(const ref_sab *sab, const double *ein_grid, int NE, const double *e_bins,
                                    int nbins, int order, double *sab_int)
{
    int L = sab_orders(order), groups = nbins - 1, iEin, g, imu, l, isab, iE, iE_lo, iE_hi, NEout, nmu = sab->n_inelastic_mu,
        NEi = sab->n_inelastic_e_in;
    double sig, mu, Ein, f, f_lo, f_hi, mult;
    double *distro = (double *)calloc((size_t)L * groups * NEi, sizeof(double));
#define DI(l, g, i) distro[((l)-1) + (size_t)L * (((g)-1) + (size_t)groups * ((i)-1))]

    memset(sab_int, 0, sizeof(double) * (size_t)L * groups * NE);

    for (iEin = 1; iEin <= NEi; ++iEin) {
        const double *Eout_arr = sab->cont_e_out + sab->cont_off[iEin - 1];
        const double *pdf_in = sab->cont_pdf + sab->cont_off[iEin - 1];
        const double *mu_arr = sab->cont_mu + (size_t)sab->cont_off[iEin - 1] * nmu;
        double *pdf;
#define MU(imu, iE) mu_arr[(size_t)((iE)-1) * nmu + ((imu)-1)]
        NEout = sab->cont_n_e_out[iEin - 1];
        pdf = (double *)malloc(sizeof(double) * NEout);
        for (iE_lo = 1; iE_lo <= NEout - 1; ++iE_lo)
            A1(pdf, iE_lo) = A1(pdf_in, iE_lo) * (A1(Eout_arr, iE_lo + 1) - A1(Eout_arr, iE_lo));
        A1(pdf, NEout) = ZERO;

        for (g = 1; g <= groups; ++g) {
            if (A1(e_bins, g) < A1(Eout_arr, 1)) {
                iE_lo = 1;
            } else if (A1(e_bins, g) >= A1(Eout_arr, NEout)) {
                for (l = 1; l <= L; ++l) DI(l, g, iEin) = ZERO;
                continue;
            } else {
                iE_lo = ref_binary_search(Eout_arr, NEout, A1(e_bins, g));
                f_lo = (A1(e_bins, g) - A1(Eout_arr, iE_lo)) / (A1(Eout_arr, iE_lo + 1) - A1(Eout_arr, iE_lo));
                mult = f_lo * A1(pdf, iE_lo);
                for (imu = 1; imu <= nmu; ++imu) {
                    mu = (ONE - f_lo) * MU(imu, iE_lo) + f_lo * MU(imu, iE_lo + 1);
                    for (l = 1; l <= L; ++l) DI(l, g, iEin) = DI(l, g, iEin) + sab_basis(L, l - 1, mu) * mult;
                }
                iE_lo = iE_lo + 1;
            }
            if (A1(e_bins, g + 1) < A1(Eout_arr, 1)) {
                for (l = 1; l <= L; ++l) DI(l, g, iEin) = ZERO;
                continue;
            } else if (A1(e_bins, g + 1) >= A1(Eout_arr, NEout)) {
                iE_hi = NEout - 1;
            } else {
                iE_hi = ref_binary_search(Eout_arr, NEout, A1(e_bins, g + 1));
                f_hi = (A1(e_bins, g + 1) - A1(Eout_arr, iE_hi)) / (A1(Eout_arr, iE_hi + 1) - A1(Eout_arr, iE_hi));
                mult = f_hi * A1(pdf, iE_hi);
                for (imu = 1; imu <= nmu; ++imu) {
                    mu = (ONE - f_hi) * MU(imu, iE_hi) + f_hi * MU(imu, iE_hi + 1);
                    for (l = 1; l <= L; ++l) DI(l, g, iEin) = DI(l, g, iEin) + sab_basis(L, l - 1, mu) * mult;
                }
                iE_hi = iE_hi - 1;
            }
            for (iE = iE_lo; iE <= iE_hi; ++iE) {
                for (imu = 1; imu <= nmu; ++imu)
                    for (l = 1; l <= L; ++l)
                        DI(l, g, iEin) = DI(l, g, iEin) + sab_basis(L, l - 1, MU(imu, iE)) * A1(pdf, iE);
            }
            for (l = 1; l <= L; ++l) DI(l, g, iEin) = DI(l, g, iEin) / (double)nmu;
        }
        free(pdf);
#undef MU
    }

    for (iEin = 1; iEin <= NE; ++iEin) {
        Ein = A1(ein_grid, iEin);
        if (Ein <= A1(sab->inelastic_e_in, 1)) {
            isab = 1;
            sig = A1(sab->inelastic_sigma, isab);
            for (g = 1; g <= groups; ++g)
                for (l = 1; l <= L; ++l) SI(l, g, iEin) = DI(l, g, isab) * sig;
        } else if (Ein >= sab->threshold_inelastic) {
            continue; /* already zero */
        } else {
            isab = ref_binary_search(sab->inelastic_e_in, NEi, Ein);
            f = (Ein - A1(sab->inelastic_e_in, isab)) /
                (A1(sab->inelastic_e_in, isab + 1) - A1(sab->inelastic_e_in, isab));
            sig = (ONE - f) * A1(sab->inelastic_sigma, isab) + f * A1(sab->inelastic_sigma, isab + 1);
            for (g = 1; g <= groups; ++g)
                for (l = 1; l <= L; ++l)
                    SI(l, g, iEin) = ((ONE - f) * DI(l, g, isab) + f * DI(l, g, isab + 1)) * sig;
        }
    }
    free(distro);
#undef DI
}

/* calc_scattsab (src/scatt.F90:543-596) + combine_sab_grid (src/sab.F90:415-454).
 * scatt_mat is (order+1, groups, NE) column-major; el_out / inel_out (nullable) receive the
 * two partial integrals.  n_threads is accepted for symmetry and ignored (serial semantics). */
int ref_sab_calc(void *h, const double *e_bins, int n_bins, int order, const double *Ein, int NE, double *scatt_mat,
                 double *el_out, double *inel_out, int n_threads)
{
    const ref_sab *sab = (const ref_sab *)h;
    int L = sab_orders(order), groups = n_bins - 1, iE, g, l;
    size_t n = (size_t)L * groups * NE;
    double *el = (double *)malloc(sizeof(double) * n), *inel = (double *)malloc(sizeof(double) * n);
    double norm_sum;
    double *sab_int;
    (void)n_threads;

    integrate_sab_el(sab, Ein, NE, e_bins, n_bins, order, el);
    if (sab->secondary_mode == REF_SAB_SECONDARY_EQUAL || sab->secondary_mode == REF_SAB_SECONDARY_SKEWED)
        integrate_sab_inel_disc(sab, Ein, NE, e_bins, n_bins, order, inel);
    else if (sab->secondary_mode == REF_SAB_SECONDARY_CONT)
        integrate_sab_inel_cont(sab, Ein, NE, e_bins, n_bins, order, inel);
    else
        memset(inel, 0, sizeof(double) * n);

    sab_int = scatt_mat;
    for (iE = 1; iE <= NE; ++iE) {
        for (g = 1; g <= groups; ++g)
            for (l = 1; l <= L; ++l) {
                size_t k = ((l)-1) + (size_t)L * (((g)-1) + (size_t)groups * ((iE)-1));
                sab_int[k] = el[k] + inel[k];
            }
        norm_sum = ZERO;
        if (g_sab_tabular) {   /* total probability = sum over groups and cosine bins */
            for (g = 1; g <= groups; ++g)
                for (l = 1; l <= L; ++l) norm_sum = norm_sum + SI(l, g, iE);
        } else {
            for (g = 1; g <= groups; ++g) norm_sum = norm_sum + SI(1, g, iE);
        }
        if (norm_sum > ZERO) {
            norm_sum = ONE / norm_sum;
            for (g = 1; g <= groups; ++g)
                for (l = 1; l <= L; ++l) SI(l, g, iE) = SI(l, g, iE) * norm_sum;
        } else {
            for (g = 1; g <= groups; ++g)
                for (l = 1; l <= L; ++l) SI(l, g, iE) = ZERO;
        }
    }
    if (NE >= 2)
        for (g = 1; g <= groups; ++g)
            for (l = 1; l <= L; ++l) SI(l, g, NE) = SI(l, g, NE - 1);

    if (el_out) memcpy(el_out, el, sizeof(double) * n);
    if (inel_out) memcpy(inel_out, inel, sizeof(double) * n);
    free(el);
    free(inel);
    return ref_error_count();
}

/* scatt_type = SCATT_TYPE_TABULAR (1): `order` cosine bins; scatt_mat is (order, groups, NE). */
int ref_sab_calc_tabular(void *h, const double *e_bins, int n_bins, int order, const double *Ein, int NE,
                         double *scatt_mat, double *el_out, double *inel_out, int n_threads)
{
    int rc;
    g_sab_tabular = 1;
    rc = ref_sab_calc(h, e_bins, n_bins, order, Ein, NE, scatt_mat, el_out, inel_out, n_threads);
    g_sab_tabular = 0;
    return rc;
}
