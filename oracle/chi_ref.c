/*
 * ORACLE (test infrastructure only): fission-spectrum (chi) integration, SURVEY 8f row N4.
 * Restates calc_chi (src/chi.F90:21-163), chi_beta / chi_prob / chi_integrate
 * (src/chidata_header.F90:143-494) and nu_total / nu_delayed (src/fission.F90:18-103).
 * Parity unpinned: the reference holds no test for these routines; pinned by analytic properties
 * (tests/test_chi.py).  Index convention as in the rest of the oracle: 1-based through A1().
 *
 * Reference behaviour kept on purpose:
 *  - law 7: a group edge above Ein-U is replaced by U, not Ein-U (:358,362);
 *  - laws 7/9/11 `return` with zeros when Ein <= U, skipping the final normalisation (:354,387,419);
 *    the laws that only warn (1,3,5,12,44,66,67) fall through to it and turn 0 into 0*(1/0) = NaN (:483-492);
 *  - law 4 reads the cdf lin-lin whatever INTT' says and extrapolates past the table (:313-320);
 *  - chi_total is overwritten by chi_prompt*(1+prob) with the LAST prompt law's prob (src/chi.F90:131);
 *  - p_valid multiplies prob only when the law has a successor and NR > 0 (:210-212).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "ndpp_oracle.h"

/* src/fission.F90:18-45 */
static double nu_total(int type, const double *d, double E)
{
    if (type == 1) { /* NU_POLYNOMIAL */
        int NC = (int)A1(d, 1), i;
        double nu = 0.0;
        for (i = 0; i <= NC - 1; ++i) nu = nu + A1(d, i + 2) * __builtin_powi(E, i);
        return nu;
    }
    if (type == 2) return ref_interpolate_tab1(d, E);
    ref_fatal("No neutron emission data for table");
    return 0.0;
}

/* src/fission.F90:90-103 */
static double nu_delayed(int type, const double *d, double E)
{
    if (type == 2) return ref_interpolate_tab1(d, E);
    return 0.0;
}

/* chi_prob, src/chidata_header.F90:158-214 */
static double chi_prob(const ref_chi_slot *s, const double *pool, int n_grid, const double *energy, const double *fission,
                       int n_precursor, const double *prec, double Ein)
{
    int NR = 0, NE = 0, lc, j;
    double f, prob;
    if (s->delayed) {
        lc = 1;
        for (j = 1; j <= n_precursor; ++j) {
            NR = (int)A1(prec, lc + 1);
            NE = (int)A1(prec, lc + 2 + 2 * NR);
            if (j == s->precursor) break;
            lc = lc + 2 + 2 * NR + 2 * NE + 1;
        }
        return ref_interpolate_tab1(prec + lc, Ein);
    }
    if (Ein < A1(energy, 1)) {
        j = 1; f = 0.0;
    } else if (Ein >= A1(energy, n_grid)) {
        j = n_grid - 1; f = 1.0;
    } else {
        j = ref_binary_search(energy, n_grid, Ein);
        f = (Ein - A1(energy, j)) / (A1(energy, j + 1) - A1(energy, j));
    }
    if (A1(energy, j) == A1(energy, j + 1)) j = j + 1;
    if (j < s->threshold) {
        prob = 0.0;
    } else {
        const double *sigma = pool + s->sigma_off;
        prob = ((1.0 - f) * A1(sigma, j - s->threshold + 1) + f * A1(sigma, j - s->threshold + 2)) /
               ((1.0 - f) * A1(fission, j) + f * A1(fission, j + 1));
    }
    if (s->use_pvalid) prob = prob * ref_interpolate_tab1(pool + s->pvalid_off, Ein);
    return prob;
}

/* chi_integrate, src/chidata_header.F90:220-494 */
static void chi_integrate(const ref_chi_slot *s, const double *pool, const double *E_bins, int groups, double Ein,
                          double *chis)
{
    const double *data = pool + s->data_off;
    int NR, NE, NP, iE, INTTp, INTT, ND, lEout_min, lc, g, histogram_interp;
    double T, U, I, x, x0, Egp1, Eg, Watt_a, Watt_b, interp, runsum;

    for (g = 1; g <= groups; ++g) A1(chis, g) = 0.0;
    switch (s->law) {
    case 4:
    case 61:
        histogram_interp = 0;
        NR = (int)A1(data, 1);
        NE = (int)A1(data, 2 + 2 * NR);
        if (NR == 1) {
            if (s->law == 4) histogram_interp = (A1(data, 3) == 1);
        } else if (NR > 1) {
            ref_fatal("Multiple interpolation regions not supported while attempting to sample continuous tabular "
                      "distribution.");
            return;
        }
        lc = 2 + 2 * NR;
        if (Ein < A1(data, lc + 1)) {
            iE = 1; x = 0.0;
        } else if (Ein >= A1(data, lc + NE)) {
            iE = NE - 1; x = 1.0;
        } else {
            iE = ref_binary_search(data + lc, NE, Ein);
            x = (Ein - A1(data, lc + iE)) / (A1(data, lc + iE + 1) - A1(data, lc + iE));
        }
        if (!histogram_interp) {
            if (x > 0.5) iE = iE + 1;
        }
        lc = (int)A1(data, 2 + 2 * NR + NE + iE);
        INTTp = (int)A1(data, lc + 1);
        NP = (int)A1(data, lc + 2);
        if (INTTp > 10) {
            INTT = INTTp % 10;
            ND = (INTTp - INTT) / 10;
        } else {
            INTT = INTTp;
            ND = 0;
        }
        if (ND > 0) {
            ref_fatal("Discrete lines in continuous tabular distributed not yet supported");
            return;
        }
        lc = lc + 3;
        lEout_min = lc;
        runsum = 0.0;
        for (g = 1; g <= groups; ++g) {
            for (iE = lEout_min; iE <= NP + lc - 2; ++iE)
                if (A1(data, iE + 1) > A1(E_bins, g + 1)) break;
            if (iE == NP + lc - 1) iE = iE - 1;
            interp = (A1(E_bins, g + 1) - A1(data, iE)) / (A1(data, iE + 1) - A1(data, iE));
            A1(chis, g) = (A1(data, iE + 2 * NP) + interp * (A1(data, iE + 1 + 2 * NP) - A1(data, iE + 2 * NP)));
            A1(chis, g) = A1(chis, g) - runsum;
            runsum = runsum + A1(chis, g);
            lEout_min = iE;
        }
        break;
    case 7:
        NR = (int)A1(data, 1);
        NE = (int)A1(data, 2 + 2 * NR);
        T = ref_interpolate_tab1(data, Ein);
        lc = 2 + 2 * NR + 2 * NE;
        U = A1(data, lc + 1);
        if (Ein - U <= 0.0) return;
        x = (Ein - U) / T;
        I = sqrt(T * T * T) * (sqrt(0.25 * REF_PI) * erf(x) - x * exp(-x));
        for (g = 1; g <= groups; ++g) {
            Egp1 = A1(E_bins, g + 1);
            if (Egp1 > Ein - U) Egp1 = U;
            A1(chis, g) = 0.5 * (sqrt(REF_PI * T) * erf(sqrt(Egp1 / T)) * exp(Egp1 / T) - 2.0 * sqrt(Egp1)) * T * exp(-Egp1 / T);
            Eg = A1(E_bins, g);
            if (Eg > Ein - U) Eg = U;
            A1(chis, g) = A1(chis, g) -
                          (0.5 * (sqrt(REF_PI * T) * erf(sqrt(Eg / T)) * exp(Eg / T) - 2.0 * sqrt(Eg)) * T * exp(-Eg / T));
            A1(chis, g) = A1(chis, g) / I;
        }
        break;
    case 9:
        NR = (int)A1(data, 1);
        NE = (int)A1(data, 2 + 2 * NR);
        T = ref_interpolate_tab1(data, Ein);
        lc = 2 + 2 * NR + 2 * NE;
        U = A1(data, lc + 1);
        x = (Ein - U) / T;
        if (Ein - U <= 0.0) return;
        for (g = 1; g <= groups; ++g) {
            Egp1 = A1(E_bins, g + 1);
            Eg = A1(E_bins, g);
            if (Egp1 > (Ein - U)) Egp1 = Ein - U;
            if (Eg > (Ein - U)) Eg = Ein - U;
            A1(chis, g) = (Egp1 * exp(x) + T * exp(x)) * exp(-Egp1 / T);
            A1(chis, g) = A1(chis, g) - (Eg * exp(x) + T * exp(x)) * exp(-Eg / T);
            A1(chis, g) = A1(chis, g) / (T * (x - exp(x) + 1.0));
        }
        break;
    case 11:
        NR = (int)A1(data, 1);
        NE = (int)A1(data, 2 + 2 * NR);
        Watt_a = ref_interpolate_tab1(data, Ein);
        lc = 2 + 2 * (NR + NE);
        Watt_b = ref_interpolate_tab1(data + lc, Ein);
        NR = (int)A1(data, lc + 1);
        NE = (int)A1(data, lc + 2 + 2 * NR);
        lc = lc + 2 + 2 * (NR + NE);
        U = A1(data, lc + 1);
        x = (Ein - U) / Watt_a;
        if (Ein - U <= 0.0) return;
        x0 = Watt_a * Watt_b * 0.25;
        I = 0.25 * sqrt(REF_PI * (Watt_a * Watt_a * Watt_a) * Watt_b) * exp(x0) *
                (erf(sqrt(x) - sqrt(x0)) + erf(sqrt(x) + sqrt(x0))) -
            Watt_a * exp(-x * sinh(Watt_a * Watt_b * x));
        Watt_b = sqrt(Watt_b);
        x = sqrt(REF_PI * Watt_a) * Watt_b * exp(0.25 * Watt_a * (Watt_b * Watt_b));
        for (g = 1; g <= groups; ++g) {
            Egp1 = A1(E_bins, g + 1);
            if (Egp1 > U) Egp1 = U;
            A1(chis, g) = (-x * erf((Watt_a * Watt_b - 2.0 * sqrt(Egp1) / (2.0 * Watt_a))) +
                           x * erf((Watt_a * Watt_b + 2.0 * sqrt(Egp1) / (2.0 * Watt_a))) -
                           2.0 * (exp(2.0 * Watt_b * sqrt(Egp1)) * exp(-(Watt_a * Watt_b * sqrt(Egp1)) / Watt_a)));
            Eg = A1(E_bins, g);
            if (Eg > U) Eg = U;
            A1(chis, g) = A1(chis, g) -
                          (-x * erf((Watt_a * Watt_b - 2.0 * sqrt(Eg) / (2.0 * Watt_a))) +
                           x * erf((Watt_a * Watt_b + 2.0 * sqrt(Eg) / (2.0 * Watt_a))) -
                           2.0 * (exp(2.0 * Watt_b * sqrt(Eg)) * exp(-(Watt_a * Watt_b * sqrt(Eg)) / Watt_a)));
            A1(chis, g) = 0.25 * Watt_a * A1(chis, g) / I;
        }
        break;
    default: /* laws 1, 3, 5, 12, 44, 66, 67: "Not Yet Supported" warning, chis stays zero */
        break;
    }
    I = 0.0;
    for (g = 1; g <= groups; ++g) I = I + A1(chis, g);
    if (I != 1.0) {
        I = 1.0 / I;
        for (g = 1; g <= groups; ++g) A1(chis, g) = A1(chis, g) * I;
    }
}

/* calc_chi's E_in loop, src/chi.F90:120-153; E_grid (the merged grid, :96-112) comes from the caller.
 * Slots: the prompt laws in the reference's order (reaction by reaction, nested laws in chain order), then the
 * delayed laws by precursor group.  Outputs: chi_total[NE][G], chi_prompt[NE][G], chi_delay[n_precursor][NE][G]. */
int ref_calc_chi(int n_grid, const double *energy, const double *fission, int nu_t_type, const double *nu_t_data,
                 int nu_d_type, const double *nu_d_data, int n_precursor, const double *precursor_data, int n_slots,
                 const ref_chi_slot *slots, const double *pool, const double *E_bins, int n_bins, const double *Ein_grid,
                 int NE, double *chi_total, double *chi_prompt, double *chi_delay)
{
    const int groups = n_bins - 1;
    int iE, i, g, n_prompt = 0;
    double *chi_p = (double *)malloc(sizeof(double) * (size_t)(groups > 0 ? groups : 1));
    ref_error_clear();
    for (i = 0; i < n_slots; ++i)
        if (!slots[i].delayed) n_prompt++;
    for (iE = 0; iE < NE; ++iE) {
        const double Ein = Ein_grid[iE];
        double *tot = chi_total + (size_t)iE * groups, *pr = chi_prompt + (size_t)iE * groups;
        double beta, prob = 0.0, norm;
        for (g = 0; g < groups; ++g) tot[g] = pr[g] = 0.0;
        beta = nu_delayed(nu_d_type, nu_d_data, Ein) / nu_total(nu_t_type, nu_t_data, Ein);
        for (i = 0; i < n_prompt; ++i) {
            chi_integrate(&slots[i], pool, E_bins, groups, Ein, chi_p);
            prob = chi_prob(&slots[i], pool, n_grid, energy, fission, n_precursor, precursor_data, Ein);
            for (g = 0; g < groups; ++g) {
                tot[g] = tot[g] + prob * (1.0 - beta) * chi_p[g];
                pr[g] = pr[g] + prob * chi_p[g];
            }
        }
        for (g = 0; g < groups; ++g) tot[g] = pr[g] + prob * pr[g];
        for (i = n_prompt; i < n_slots; ++i) {
            double *dl = chi_delay + ((size_t)(i - n_prompt) * NE + iE) * groups;
            chi_integrate(&slots[i], pool, E_bins, groups, Ein, dl);
            prob = chi_prob(&slots[i], pool, n_grid, energy, fission, n_precursor, precursor_data, Ein);
            for (g = 0; g < groups; ++g) tot[g] = tot[g] + prob * beta * dl[g];
        }
        norm = 0.0;
        for (g = 0; g < groups; ++g) norm = norm + tot[g];
        if (norm > 0.0)
            for (g = 0; g < groups; ++g) tot[g] = tot[g] / norm;
        norm = 0.0;
        for (g = 0; g < groups; ++g) norm = norm + pr[g];
        if (norm > 0.0)
            for (g = 0; g < groups; ++g) pr[g] = pr[g] / norm;
        for (i = n_prompt; i < n_slots; ++i) {
            double *dl = chi_delay + ((size_t)(i - n_prompt) * NE + iE) * groups;
            norm = 0.0;
            for (g = 0; g < groups; ++g) norm = norm + dl[g];
            if (norm > 0.0)
                for (g = 0; g < groups; ++g) dl[g] = dl[g] / norm;
        }
    }
    free(chi_p);
    return ref_error_count();
}
