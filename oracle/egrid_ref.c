/*
 * ORACLE (test infrastructure only): the incoming-energy grid builders, restated line by line.
 *   create_Ein_grid      src/scatt.F90:166-236
 *   combine_Eins         src/scatt.F90:246-299
 *   add_elastic_Eins     src/scatt.F90:311-419
 *   add_one_more_point   src/scatt.F90:426-445
 *   add_inelastic_Eins   src/scatt.F90:456-536
 *   sab_egrid            src/sab.F90:460-568
 * (merge: scattdata_ref.c / src/array_merge.F90).  log, exp and sqrt are the C library's, as in the gfortran build.
 * Parity unpinned: the reference holds no test for these routines; the restatement is held by the independent numpy
 * version of the same text (ndpp_b200/egrid.py, written first and separately) in tests/test_egrid.py.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "ndpp_oracle.h"

#define ZERO 0.0
#define ONE 1.0
#define TWO 2.0

void ref_fatal(const char *msg);

/* allocatable real(8) array */
typedef struct {
    double *v;
    int n;
} darr;

static void darr_set(darr *a, const double *src, int n)
{
    double *p = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    if (n > 0) memcpy(p, src, sizeof(double) * (size_t)n);
    free(a->v);
    a->v = p;
    a->n = n;
}

/* call merge(a, b, result): result may alias neither input here */
static void merge_into(const double *a, int na, const double *b, int nb, darr *result)
{
    double *tmp = (double *)malloc(sizeof(double) * (size_t)(na + nb + 1));
    int n;
    /* a zero-size a(:): the Fortran text reads a(0), undefined; every defined reading leaves b unchanged */
    if (na == 0) { memcpy(tmp, b, sizeof(double) * (size_t)nb); n = nb; }
    else n = ref_merge(a, na, b, nb, tmp);
    darr_set(result, tmp, n);
    free(tmp);
}

/* src/scatt.F90:426-445; the literal 1.0E-3 is single precision */
static void add_one_more_point(darr *Ein)
{
    double *t = (double *)malloc(sizeof(double) * (size_t)(Ein->n + 1));
    double Ehi = A1(Ein->v, Ein->n);
    memcpy(t, Ein->v, sizeof(double) * (size_t)Ein->n);
    t[Ein->n] = Ehi * (ONE + (double)1.0E-3f);
    darr_set(Ein, t, Ein->n + 1);
    free(t);
}

/* src/scatt.F90:311-419 */
static void add_elastic_Eins(double awr, double kT, double cutoff, const double *E_bins, int nb, int extend_pts, darr *Ein)
{
    darr old_grid = {0, 0};
    double *new_pts;
    double alpha, lo_shift, Ehi, Elo, newE, dElo, dEhi;
    int num_pts, g, i;

    darr_set(&old_grid, Ein->v, Ein->n);
    alpha = ((awr - ONE) / (awr + ONE)) * ((awr - ONE) / (awr + ONE));
    new_pts = (double *)calloc((size_t)extend_pts * (size_t)nb + 1, sizeof(double));
    lo_shift = TWO * kT * (awr + ONE) / awr;
    num_pts = 0;
    if (cutoff != ZERO) {
        for (g = 1; g <= nb - 1; ++g) {
            Ehi = A1(E_bins, g + 1);
            Elo = A1(E_bins, g);
            if (Ehi <= cutoff) {
                if (Ehi - lo_shift > Elo)
                    dElo = log(Ehi / (Ehi - lo_shift)) / (double)extend_pts;
                else
                    dElo = log(Ehi / 1E-11) / (double)extend_pts;
                for (i = -extend_pts; i <= -1; ++i) {
                    newE = Ehi * exp((double)i * dElo);
                    if (newE >= Elo) {
                        num_pts = num_pts + 1;
                        A1(new_pts, num_pts) = newE;
                    }
                }
            } else if (Elo < cutoff) {
                Ehi = cutoff;
                dElo = log(Ehi / (Ehi - lo_shift)) / (double)extend_pts;
                for (i = -extend_pts; i <= -1; ++i) {
                    newE = Ehi * exp((double)i * dElo);
                    if (newE > Elo) {
                        num_pts = num_pts + 1;
                        A1(new_pts, num_pts) = newE;
                    }
                }
            }
        }
        merge_into(new_pts, num_pts, old_grid.v, old_grid.n, Ein);
        darr_set(&old_grid, Ein->v, Ein->n);
        memset(new_pts, 0, sizeof(double) * ((size_t)extend_pts * (size_t)nb + 1));
        num_pts = 0;
    }
    dEhi = 7.0 * log(ONE / alpha) / (double)extend_pts;
    for (g = 1; g <= nb - 1; ++g) {
        if (A1(E_bins, g) == ZERO) continue;
        Ehi = A1(E_bins, g + 1);
        for (i = 1; i <= extend_pts - 1; ++i) {
            newE = A1(E_bins, g) * exp((double)i * dEhi);
            if (newE < Ehi) {
                num_pts = num_pts + 1;
                A1(new_pts, num_pts) = newE;
            } else
                break;
        }
    }
    merge_into(new_pts, num_pts, old_grid.v, old_grid.n, Ein);
    free(new_pts);
    free(old_grid.v);
}

/* src/scatt.F90:456-536 */
static void add_inelastic_Eins(int n_slots, const int *is_init, const double *Q_value, double awr, const double *E_bins,
                               int nb, double thresh, int inel_extend_pts, darr *Ein)
{
    int i_rxn, g, i, num_pts;
    double Ef, D, Fp, Fm, Ecp, Ecm, Eg, Q, dE, Ehi, Elo;
    double *new_pts = (double *)malloc(sizeof(double) * ((size_t)inel_extend_pts * (size_t)nb + 1));
    darr old_grid = {0, 0};

    for (i_rxn = 1; i_rxn <= n_slots; ++i_rxn) {
        if (!A1(is_init, i_rxn)) continue;
        Q = -A1(Q_value, i_rxn);
        if (Q != ZERO) {
            for (g = 2; g <= nb - 1; ++g) {
                darr_set(&old_grid, Ein->v, Ein->n);
                num_pts = 0;
                Eg = A1(E_bins, g);
                Ef = (ONE + awr) / (awr)*Eg;
                D = ((awr * awr) * (ONE + Ef / Q) - ONE) * (Ef / Q);
                Fp = (ONE + sqrt(D)) / (ONE + Ef / Q);
                Fm = (ONE - sqrt(D)) / (ONE + Ef / Q);
                Ecp = ((ONE + awr) / (awr)*Q) / (ONE - Fp * Fp / (awr * awr));
                Ecm = ((ONE + awr) / (awr)*Q) / (ONE - Fm * Fm / (awr * awr));
                if (Ecp > Ecm) {
                    Elo = Ecm;
                    Ehi = Ecp;
                } else {
                    Elo = Ecp;
                    Ehi = Ecm;
                }
                if (Elo < thresh) Elo = thresh;
                if (Ehi < thresh) Ehi = thresh;
                if (Elo != Ehi) {
                    dE = log(Ehi / Elo) / (double)inel_extend_pts;
                    for (i = 1; i <= inel_extend_pts - 1; ++i) {
                        num_pts = num_pts + 1;
                        A1(new_pts, num_pts) = Elo * exp((double)i * dE);
                    }
                    merge_into(new_pts, num_pts, old_grid.v, old_grid.n, Ein);
                }
            }
        }
    }
    free(new_pts);
    free(old_grid.v);
}

/* src/scatt.F90:246-299 */
static void combine_Eins(int n_slots, const int *is_init, const int *MT, const double *const *E_grid, const int *NE,
                         const double *E_bins, int nb, darr *Ein, int *only_el)
{
    darr new_grid = {0, 0}, tmp_grid = {0, 0};
    int i_rxn, iEmax;
    double max_grp, min_grp;

    *only_el = 1;
    darr_set(&new_grid, Ein->v, 1);
    for (i_rxn = 1; i_rxn <= n_slots; ++i_rxn) {
        const double *eg = A1(E_grid, i_rxn);
        const int ne = A1(NE, i_rxn);
        if (!A1(is_init, i_rxn)) continue;
        if (A1(MT, i_rxn) != REF_ELASTIC) *only_el = 0;
        min_grp = A1(E_bins, 1);
        max_grp = A1(E_bins, nb);
        if (min_grp >= A1(eg, ne))
            continue;
        else if (max_grp <= A1(eg, 1))
            continue;
        else {
            if (max_grp >= A1(eg, ne))
                iEmax = ne;
            else
                iEmax = ref_binary_search(eg, ne, max_grp);
        }
        merge_into(eg, iEmax, new_grid.v, new_grid.n, &tmp_grid);
        darr_set(&new_grid, tmp_grid.v, tmp_grid.n);
    }
    darr_set(&tmp_grid, Ein->v, Ein->n);
    merge_into(new_grid.v, new_grid.n, tmp_grid.v, tmp_grid.n, Ein);
    free(tmp_grid.v);
    free(new_grid.v);
}

/* src/scatt.F90:166-236.  Slots as calc_scatt holds them (rxn_data(:)): is_init, MT and Q_value of the slot's reaction,
 * its E_grid.  Ein_el / Ein_inel: caller's buffers of cap doubles; *n_inel = 0 for a nuclide with elastic scattering only.
 * Returns 0, or 1 when a buffer is too small (the counts are set either way). */
int ref_create_ein_grid(int n_slots, const int *is_init, const int *MT, const double *Q_value, const double *const *E_grid,
                        const int *NE, const double *E_bins, int nb, const double *nuc_grid, int n_grid, double awr,
                        double kT, double cutoff, double thresh, int extend_pts, int inel_extend_pts, double *Ein_el,
                        int *n_el, double *Ein_inel, int *n_inel, int cap)
{
    darr el = {0, 0}, inel = {0, 0};
    int iEmax, iEthresh, only_el, rc = 0;

    if (A1(E_bins, nb) >= A1(nuc_grid, n_grid))
        iEmax = n_grid;
    else
        iEmax = ref_binary_search(nuc_grid, n_grid, A1(E_bins, nb));
    merge_into(nuc_grid, iEmax, E_bins, nb, &el);
    combine_Eins(n_slots, is_init, MT, E_grid, NE, E_bins, nb, &el, &only_el);
    add_elastic_Eins(awr, kT, cutoff, E_bins, nb, extend_pts, &el);
    add_one_more_point(&el);
    *n_inel = 0;
    if (!only_el) {
        iEthresh = ref_binary_search(el.v, el.n, thresh);
        darr_set(&inel, el.v + (iEthresh - 1), el.n - iEthresh + 1);
        add_inelastic_Eins(n_slots, is_init, Q_value, awr, E_bins, nb, thresh, inel_extend_pts, &inel);
        iEthresh = ref_binary_search(inel.v, inel.n, A1(E_bins, nb));
        inel.n = iEthresh;
        add_one_more_point(&inel);
        *n_inel = inel.n;
        if (inel.n <= cap) memcpy(Ein_inel, inel.v, sizeof(double) * (size_t)inel.n); else rc = 1;
    }
    *n_el = el.n;
    if (el.n <= cap) memcpy(Ein_el, el.v, sizeof(double) * (size_t)el.n); else rc = 1;
    free(el.v);
    free(inel.v);
    return rc;
}

/* src/sab.F90:460-568.  inelastic_e_out(j, i) = e_out[(i-1) * n_e_out + (j-1)] (Fortran column-major (j, i)).
 * elastic_e_in may be NULL (not allocated).  Returns the number of points, or -n when cap is too small. */
int ref_sab_egrid(const double *inelastic_e_in, int n_in, const double *elastic_e_in, int n_el, const double *inelastic_e_out,
                  int n_e_out, int secondary_mode, const double *energy_bins, int nb, int sab_epts_per_bin, int extend_pts,
                  double *Ein_out, int cap)
{
    darr Ein = {0, 0}, t = {0, 0}, t2 = {0, 0};
    double max_ein, dE, Eo1, Eo2, Ei1, Ei2;
    int i, j, g, g1, g2, num_pts, i_max_ein, iE, n;

    if (elastic_e_in) {
        merge_into(inelastic_e_in, n_in, elastic_e_in, n_el, &t);
        merge_into(t.v, t.n, energy_bins, nb, &Ein);
        max_ein = A1(inelastic_e_in, n_in) > A1(elastic_e_in, n_el) ? A1(inelastic_e_in, n_in) : A1(elastic_e_in, n_el);
    } else {
        merge_into(inelastic_e_in, n_in, energy_bins, nb, &Ein);
        max_ein = A1(inelastic_e_in, n_in);
    }
    if (secondary_mode != REF_SAB_SECONDARY_CONT) {
        double *pts = (double *)malloc(sizeof(double) * (size_t)(nb + 1));
        for (i = 1; i <= n_in - 1; ++i) {
            Ei1 = A1(inelastic_e_in, i);
            Ei2 = A1(inelastic_e_in, i + 1);
            for (j = 1; j <= n_e_out; ++j) {
                num_pts = 0;
                Eo1 = inelastic_e_out[(size_t)(i - 1) * n_e_out + (j - 1)];
                g1 = ref_binary_search(energy_bins, nb, Eo1);
                Eo2 = inelastic_e_out[(size_t)i * n_e_out + (j - 1)];
                g2 = ref_binary_search(energy_bins, nb, Eo2);
                if (Eo2 < Eo1) {
                    g = g1;
                    g2 = g1;
                    g1 = g;
                }
                for (g = g1 + 1; g <= g2; ++g) {
                    num_pts = num_pts + 1;
                    A1(pts, num_pts) = (A1(energy_bins, g) - Eo1) / (Eo2 - Eo1) * (Ei2 - Ei1) + Ei1;
                }
                if (num_pts > 0) {
                    merge_into(pts, num_pts, Ein.v, Ein.n, &t2);
                    darr_set(&Ein, t2.v, t2.n);
                }
            }
        }
        free(pts);
    }
    i_max_ein = ref_binary_search(Ein.v, Ein.n, max_ein);
    darr_set(&t, Ein.v, Ein.n);
    if (sab_epts_per_bin == 0) {
        n = i_max_ein;
        if (n <= cap) memcpy(Ein_out, t.v, sizeof(double) * (size_t)n);
    } else {
        n = (i_max_ein - 1) * extend_pts + i_max_ein;
        if (n <= cap) {
            j = 0;
            for (iE = 1; iE <= i_max_ein - 1; ++iE) {
                dE = (log(A1(t.v, iE + 1) / A1(t.v, iE))) / (double)(extend_pts + 1);
                j = j + 1;
                A1(Ein_out, j) = A1(t.v, iE);
                for (i = 1; i <= extend_pts; ++i) {
                    j = j + 1;
                    A1(Ein_out, j) = A1(Ein_out, j - 1) * exp(dE);
                }
            }
            A1(Ein_out, n) = A1(t.v, i_max_ein);
        }
    }
    free(Ein.v);
    free(t.v);
    free(t2.v);
    return n <= cap ? n : -n;
}
