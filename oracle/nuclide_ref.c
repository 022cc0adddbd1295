/*
 * ORACLE (test infrastructure only -- never linked into or called by the product path).
 *
 * CPU restatement of the nuclide-level part of the reference's scattering integrator:
 *   scatt_init            src/scattdata_header.F90:78-271
 *   scatt_convert_distro  src/scattdata_header.F90:325-382
 *   scatt_interp_distro   src/scattdata_header.F90:391-499
 *   integrate_distro      src/scattdata_header.F90:513-662
 *   unitbase              src/scattdata_header.F90:1521-1546
 *   calc_elastic_grid     src/scatt.F90:603-675
 *   calc_inelastic_grid   src/scatt.F90:682-778
 * A "slot" is one ScattData object, i.e. one (reaction, energy distribution) pair in the order
 * calc_scatt builds rxn_data(:) (src/scatt.F90:84-105).  The E_in loops may run with OpenMP
 * (schedule(dynamic,100) as src/scatt.F90:631,722); every column is independent, and the
 * top-of-grid column copy (:669,:770) is done after the loop so it cannot race.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "ndpp_oracle.h"

#define ZERO 0.0
#define ONE 1.0

void ref_fatal(const char *msg);
void ref_freegas_counters_flush(void);

typedef struct {
    int MT, multiplicity, threshold, scatter_in_cm, has_angle_dist, has_energy_dist, n_sigma;
    double Q;
    double *yield_tab1; /* NULL unless multiplicity_with_E */
    double *sigma;
    /* DistAngle */
    int ad_n, n_ad_data;
    double *ad_energy, *ad_data;
    int *ad_type, *ad_loc;
} ref_rxn;

typedef struct {
    int is_init, NE, order, groups, law, has_adist, has_edist, M;
    double *E_grid;
    double **distro, **Eouts, **pdfs, **cdfs;
    int *NP, *INTT;
    double awr, kT, freegas_cutoff;
    ref_rxn *rxn;
    double *edist_data;
    int n_edist_data, edist_law;
    double *p_valid_tab1;
} ref_slot;

typedef struct {
    double awr, kT, freegas_cutoff;
    int n_grid, n_bins, n_rxn, n_slots, cap_rxn, cap_slots;
    double *energy, *elastic, *e_bins, *mu;
    ref_params p;
    ref_rxn **rxns;
    int *rxn_ids;
    ref_slot *slots;
} ref_nuclide;

static double *dupd(const double *src, size_t n)
{
    double *p;
    if (!src || n == 0) return NULL;
    p = (double *)malloc(n * sizeof(double));
    memcpy(p, src, n * sizeof(double));
    return p;
}
static int *dupi(const int *src, size_t n)
{
    int *p;
    if (!src || n == 0) return NULL;
    p = (int *)malloc(n * sizeof(int));
    memcpy(p, src, n * sizeof(int));
    return p;
}

void *ref_nuclide_create(double awr, double kT, double freegas_cutoff, int n_grid, const double *energy,
                         const double *elastic_xs, const double *e_bins, int n_bins, const ref_params *p)
{
    ref_nuclide *n = (ref_nuclide *)calloc(1, sizeof(ref_nuclide));
    int i, M = p->mu_bins;
    double dmu;
    n->awr = awr;
    n->kT = kT;
    n->freegas_cutoff = freegas_cutoff;
    n->n_grid = n_grid;
    n->n_bins = n_bins;
    n->energy = dupd(energy, n_grid);
    n->elastic = dupd(elastic_xs, n_grid);
    n->e_bins = dupd(e_bins, n_bins);
    n->p = *p;
    /* scattdata_header.F90:250-257 */
    n->mu = (double *)malloc(sizeof(double) * M);
    dmu = 2.0 / (double)(M - 1);
    for (i = 1; i <= M - 1; ++i) A1(n->mu, i) = -ONE + (double)(i - 1) * dmu;
    A1(n->mu, M) = ONE;
    return n;
}

/* src/scattdata_header.F90:1502-1515 */
static int is_valid_scatter(int MT)
{
    if ((MT == REF_ELASTIC) || ((MT >= 11) && (MT <= 91)))
        if (MT != 18 && MT != 19 && MT != 20 && MT != 21 && MT != 38) return 1;
    return 0;
}

static void synth_isotropic_adist(ref_nuclide *nuc, ref_rxn *rxn)
{
    /* scattdata_header.F90:162-185 / :196-213 */
    free(rxn->ad_energy); free(rxn->ad_type); free(rxn->ad_loc); free(rxn->ad_data);
    rxn->ad_n = 2;
    rxn->ad_energy = (double *)malloc(2 * sizeof(double));
    rxn->ad_energy[1] = A1(nuc->e_bins, nuc->n_bins);
    if (A1(nuc->energy, rxn->threshold) > A1(nuc->e_bins, 1))
        rxn->ad_energy[0] = A1(nuc->energy, rxn->threshold);
    else
        rxn->ad_energy[0] = A1(nuc->e_bins, 1);
    rxn->ad_type = (int *)malloc(2 * sizeof(int));
    rxn->ad_type[0] = rxn->ad_type[1] = REF_ANGLE_ISOTROPIC;
    rxn->ad_loc = (int *)calloc(2, sizeof(int));
    rxn->ad_data = (double *)calloc(2, sizeof(double));
    rxn->n_ad_data = 2;
}

/* src/scattdata_header.F90:78-271 */
static void scatt_init(ref_nuclide *nuc, ref_slot *s, ref_rxn *rxn, int edist_assoc, int edist_law)
{
    int i, NR, lc, NP, M = nuc->p.mu_bins;

    s->is_init = 0;
    s->M = M;
    if (!is_valid_scatter(rxn->MT)) return;
    if (edist_assoc)
        if ((edist_law != 3) && (edist_law != 44) && (edist_law != 61) && (edist_law != 9) && (edist_law != 4)) return;

    if (nuc->p.scatt_type == REF_SCATT_TYPE_LEGENDRE)
        s->order = nuc->p.order + 1;
    else
        s->order = nuc->p.order;
    s->rxn = rxn;
    s->awr = nuc->awr;
    s->kT = nuc->kT;
    s->freegas_cutoff = ZERO;
    if (rxn->MT == REF_ELASTIC) s->freegas_cutoff = nuc->freegas_cutoff;

    if (rxn->has_angle_dist) {
        s->has_adist = 1;
        if (edist_assoc) {
            s->has_edist = (edist_law == 3) ? 0 : 1;
            s->law = edist_law;
        } else {
            s->has_edist = 0;
            s->law = 0;
        }
    } else if (edist_assoc) {
        if ((edist_law == 4) || (edist_law == 3) || (edist_law == 9)) {
            s->has_edist = ((edist_law == 9) || (edist_law == 4)) ? 1 : 0;
            synth_isotropic_adist(nuc, rxn);
            s->has_adist = 1;
            rxn->has_angle_dist = 1;
        } else {
            s->has_adist = 0;
            s->has_edist = 1;
        }
        s->law = edist_law;
    } else {
        synth_isotropic_adist(nuc, rxn);
        s->has_adist = 1;
        rxn->scatter_in_cm = 1;
        s->has_edist = 0;
        s->law = 0;
    }

    if (s->has_adist && !s->has_edist) {
        s->NE = rxn->ad_n;
        s->E_grid = dupd(rxn->ad_energy, s->NE);
        s->distro = (double **)calloc(s->NE, sizeof(double *));
        s->NP = (int *)calloc(s->NE, sizeof(int));
        for (i = 0; i < s->NE; ++i) {
            s->NP[i] = 1;
            s->distro[i] = (double *)calloc((size_t)M, sizeof(double));
        }
    } else if (s->has_edist && edist_law != 3) {
        const double *data = s->edist_data;
        NR = (int)A1(data, 1);
        s->NE = (int)A1(data, 2 + 2 * NR);
        lc = 2 + 2 * NR;
        s->E_grid = dupd(data + lc, s->NE);
        s->distro = (double **)calloc(s->NE, sizeof(double *));
        s->NP = (int *)calloc(s->NE, sizeof(int));
        for (i = 1; i <= s->NE; ++i) {
            lc = (int)A1(data, 2 + 2 * NR + s->NE + i);
            NP = (int)A1(data, lc + 2);
            s->NP[i - 1] = NP;
            s->distro[i - 1] = (double *)calloc((size_t)M * NP, sizeof(double));
        }
    }
    s->Eouts = (double **)calloc(s->NE, sizeof(double *));
    s->pdfs = (double **)calloc(s->NE, sizeof(double *));
    s->cdfs = (double **)calloc(s->NE, sizeof(double *));
    s->INTT = (int *)calloc(s->NE, sizeof(int));
    s->groups = nuc->n_bins - 1;
    s->is_init = 1;
}

int ref_nuclide_add_reaction(void *h, int rxn_index, int MT, double Q, int threshold, int scatter_in_cm,
                             int has_angle_dist, int has_energy_dist, int law, int multiplicity,
                             const double *yield_tab1, int n_yield, const double *sigma, int n_sigma,
                             const double *p_valid_tab1, int n_pvalid, const double *adist_energy,
                             const int *adist_type, const int *adist_loc, int n_adist_e, const double *adist_data,
                             int n_adist_data, const double *edist_data, int n_edist_data)
{
    ref_nuclide *nuc = (ref_nuclide *)h;
    ref_rxn *rxn = NULL;
    ref_slot *s;
    int i;

    for (i = 0; i < nuc->n_rxn; ++i)
        if (nuc->rxn_ids[i] == rxn_index) rxn = nuc->rxns[i];
    if (!rxn) {
        if (nuc->n_rxn == nuc->cap_rxn) {
            nuc->cap_rxn = nuc->cap_rxn ? 2 * nuc->cap_rxn : 16;
            nuc->rxns = (ref_rxn **)realloc(nuc->rxns, sizeof(ref_rxn *) * nuc->cap_rxn);
            nuc->rxn_ids = (int *)realloc(nuc->rxn_ids, sizeof(int) * nuc->cap_rxn);
        }
        rxn = (ref_rxn *)calloc(1, sizeof(ref_rxn));
        rxn->MT = MT;
        rxn->Q = Q;
        rxn->multiplicity = multiplicity;
        rxn->yield_tab1 = dupd(yield_tab1, n_yield);
        rxn->threshold = threshold;
        rxn->scatter_in_cm = scatter_in_cm;
        rxn->has_angle_dist = has_angle_dist;
        rxn->has_energy_dist = has_energy_dist;
        rxn->sigma = dupd(sigma, n_sigma);
        rxn->n_sigma = n_sigma;
        if (has_angle_dist) {
            rxn->ad_n = n_adist_e;
            rxn->ad_energy = dupd(adist_energy, n_adist_e);
            rxn->ad_type = dupi(adist_type, n_adist_e);
            rxn->ad_loc = dupi(adist_loc, n_adist_e);
            rxn->ad_data = dupd(adist_data, n_adist_data);
            rxn->n_ad_data = n_adist_data;
        }
        nuc->rxns[nuc->n_rxn] = rxn;
        nuc->rxn_ids[nuc->n_rxn] = rxn_index;
        nuc->n_rxn++;
    }
    if (nuc->n_slots == nuc->cap_slots) {
        nuc->cap_slots = nuc->cap_slots ? 2 * nuc->cap_slots : 16;
        nuc->slots = (ref_slot *)realloc(nuc->slots, sizeof(ref_slot) * nuc->cap_slots);
    }
    s = &nuc->slots[nuc->n_slots++];
    memset(s, 0, sizeof(*s));
    s->edist_data = dupd(edist_data, n_edist_data);
    s->n_edist_data = n_edist_data;
    s->edist_law = law;
    s->p_valid_tab1 = dupd(p_valid_tab1, n_pvalid);
    scatt_init(nuc, s, rxn, has_energy_dist, law);
    return 0;
}

int ref_nuclide_n_slots(void *h) { return ((ref_nuclide *)h)->n_slots; }

int ref_nuclide_slot_info(void *h, int slot, int *info)
{
    ref_nuclide *nuc = (ref_nuclide *)h;
    ref_slot *s;
    if (slot < 0 || slot >= nuc->n_slots) return 1;
    s = &nuc->slots[slot];
    info[0] = s->is_init;
    info[1] = s->NE;
    info[2] = s->law;
    info[3] = s->has_adist;
    info[4] = s->has_edist;
    info[5] = s->order;
    info[6] = s->groups;
    info[7] = s->rxn ? s->rxn->MT : 0;
    return 0;
}

int ref_nuclide_slot_row_np(void *h, int slot, int iE)
{
    ref_slot *s = &((ref_nuclide *)h)->slots[slot];
    if (!s->is_init || iE < 1 || iE > s->NE) return -1;
    return s->NP[iE - 1];
}

int ref_nuclide_slot_egrid(void *h, int slot, double *E_grid)
{
    ref_slot *s = &((ref_nuclide *)h)->slots[slot];
    if (!s->is_init) return -1;
    memcpy(E_grid, s->E_grid, sizeof(double) * s->NE);
    return s->NE;
}

/* src/scattdata_header.F90:325-382 */
static void convert_slot(ref_nuclide *nuc, ref_slot *s)
{
    int iE, iEout, iEadist, NPo, M = s->M;
    ref_rxn *rxn = s->rxn;

    if (!s->is_init) return;
    if (!s->has_edist && !s->has_adist) {
        ref_fatal("No distribution associated with this ScattData object.");
        return;
    }
    for (iE = 1; iE <= s->NE; ++iE) {
        double *d = s->distro[iE - 1];
        int NP = s->NP[iE - 1];
        memset(d, 0, sizeof(double) * (size_t)M * NP);
        if ((s->law == 0) || (s->law == 3) || (s->law == 9)) {
            ref_convert_file4(iE, nuc->mu, M, rxn->ad_energy, rxn->ad_type, rxn->ad_loc, rxn->ad_data, d);
            /* convert_file4 tail :754-760 */
            if (!s->Eouts[iE - 1]) {
                s->Eouts[iE - 1] = (double *)malloc(2 * sizeof(double));
                s->Eouts[iE - 1][0] = ZERO;
                s->Eouts[iE - 1][1] = REF_INFINITY;
                s->INTT[iE - 1] = REF_HISTOGRAM;
            }
        } else if (s->law == 4) {
            if (A1(s->E_grid, iE) <= A1(rxn->ad_energy, 1))
                iEadist = 1;
            else if (A1(s->E_grid, iE) >= A1(rxn->ad_energy, rxn->ad_n))
                iEadist = rxn->ad_n;
            else
                iEadist = ref_binary_search(rxn->ad_energy, rxn->ad_n, A1(s->E_grid, iE));
            ref_convert_file4(iEadist, nuc->mu, M, rxn->ad_energy, rxn->ad_type, rxn->ad_loc, rxn->ad_data, d);
            /* :366-368 -- size(Eouts) is 2 here (set by convert_file4's tail), so only columns
             * 1..2 receive the copy; convert_file6 law 4 then leaves the rest of distro at zero. */
            for (iEout = 2; iEout <= 2 && iEout <= NP; ++iEout) memcpy(d + (size_t)(iEout - 1) * M, d, sizeof(double) * M);
            free(s->Eouts[iE - 1]);
            s->Eouts[iE - 1] = (double *)malloc(sizeof(double) * NP);
            s->pdfs[iE - 1] = (double *)malloc(sizeof(double) * NP);
            s->cdfs[iE - 1] = (double *)malloc(sizeof(double) * NP);
            ref_convert_file6(iE, nuc->mu, M, s->edist_law, s->edist_data, &s->INTT[iE - 1], &NPo, s->Eouts[iE - 1],
                              s->pdfs[iE - 1], s->cdfs[iE - 1], d);
        } else {
            free(s->Eouts[iE - 1]);
            s->Eouts[iE - 1] = (double *)malloc(sizeof(double) * NP);
            s->pdfs[iE - 1] = (double *)malloc(sizeof(double) * NP);
            s->cdfs[iE - 1] = (double *)malloc(sizeof(double) * NP);
            ref_convert_file6(iE, nuc->mu, M, s->edist_law, s->edist_data, &s->INTT[iE - 1], &NPo, s->Eouts[iE - 1],
                              s->pdfs[iE - 1], s->cdfs[iE - 1], d);
        }
    }
}

int ref_nuclide_convert_distro(void *h)
{
    ref_nuclide *nuc = (ref_nuclide *)h;
    int i;
    for (i = 0; i < nuc->n_slots; ++i) convert_slot(nuc, &nuc->slots[i]);
    return ref_error_count();
}

int ref_nuclide_get_table(void *h, int slot, int iE, double *distro, double *Eouts, double *pdf, double *cdf, int *INTT)
{
    ref_slot *s = &((ref_nuclide *)h)->slots[slot];
    int NP;
    if (!s->is_init || iE < 1 || iE > s->NE) return 1;
    NP = s->NP[iE - 1];
    if (distro) memcpy(distro, s->distro[iE - 1], sizeof(double) * (size_t)s->M * NP);
    if (s->has_edist && s->law != 9) {
        if (Eouts && s->Eouts[iE - 1]) memcpy(Eouts, s->Eouts[iE - 1], sizeof(double) * NP);
        if (pdf && s->pdfs[iE - 1]) memcpy(pdf, s->pdfs[iE - 1], sizeof(double) * NP);
        if (cdf && s->cdfs[iE - 1]) memcpy(cdf, s->cdfs[iE - 1], sizeof(double) * NP);
    } else if (Eouts && s->Eouts[iE - 1]) {
        Eouts[0] = s->Eouts[iE - 1][0];
    }
    if (INTT) *INTT = s->INTT[iE - 1];
    return 0;
}

int ref_nuclide_set_table(void *h, int slot, int iE, const double *distro)
{
    ref_slot *s = &((ref_nuclide *)h)->slots[slot];
    if (!s->is_init || iE < 1 || iE > s->NE) return 1;
    memcpy(s->distro[iE - 1], distro, sizeof(double) * (size_t)s->M * s->NP[iE - 1]);
    return 0;
}

/* src/scattdata_header.F90:1521-1546 + the file-6 integrators */
static void unitbase_then(ref_nuclide *nuc, ref_slot *s, double Ein, int iE, int cm, double *result)
{
    int n1 = s->NP[iE - 1], n2 = s->NP[iE], M = s->M, nub1, nub2, nu, INTT;
    double *ub1 = (double *)malloc(sizeof(double) * n1), *ub2 = (double *)malloc(sizeof(double) * n2);
    double *Eout = (double *)malloc(sizeof(double) * (n1 + n2)), *pdf = (double *)malloc(sizeof(double) * (n1 + n2));
    double *fEmu = (double *)malloc(sizeof(double) * (size_t)M * (n1 + n2));

    nub1 = ref_cast_to_unitbase(s->Eouts[iE - 1], n1, ub1);
    nub2 = ref_cast_to_unitbase(s->Eouts[iE], n2, ub2);
    nu = ref_interp_unitbase(Ein, ub1, nub1, s->Eouts[iE - 1], n1, s->pdfs[iE - 1], s->INTT[iE - 1], s->distro[iE - 1],
                             A1(s->E_grid, iE), ub2, nub2, s->Eouts[iE], n2, s->pdfs[iE], s->INTT[iE], s->distro[iE],
                             A1(s->E_grid, iE + 1), M, Eout, pdf, &INTT, fEmu);
    if (cm)
        ref_integrate_file6_cm_leg(fEmu, nuc->mu, M, Ein, s->awr, Eout, nu, INTT, pdf, nuc->e_bins, nuc->n_bins,
                                   s->order, nuc->p.ne_per_grp, result);
    else
        ref_integrate_file6_lab_leg(fEmu, nuc->mu, M, Eout, nu, INTT, pdf, nuc->e_bins, nuc->n_bins, s->order, result);
    free(ub1); free(ub2); free(Eout); free(pdf); free(fEmu);
}

/* src/scattdata_header.F90:513-662; result (order x groups) must be zeroed by the caller */
static void integrate_distro(ref_nuclide *nuc, ref_slot *s, double Ein, int iE, double *result)
{
    int n = s->order * s->groups, k, pass;
    double f;
    double *distro_int;

    if (nuc->p.scatt_type != REF_SCATT_TYPE_LEGENDRE) return; /* :658 TABULAR is empty */

    if (s->has_adist && !s->has_edist) {
        f = (Ein - A1(s->E_grid, iE)) / (A1(s->E_grid, iE + 1) - A1(s->E_grid, iE));
        distro_int = (double *)malloc(sizeof(double) * n);
        for (pass = 0; pass < 2; ++pass) {
            const double *row = s->distro[iE - 1 + pass];
            for (k = 0; k < n; ++k) distro_int[k] = ZERO;
            if ((Ein < s->freegas_cutoff) && (s->rxn->MT == REF_ELASTIC)) {
                ref_integrate_freegas_leg(Ein, s->awr, s->kT, row, nuc->mu, s->M, nuc->e_bins, nuc->n_bins, s->order,
                                          &nuc->p, distro_int);
            } else if (s->rxn->scatter_in_cm) {
                ref_integrate_file4_cm_leg(row, Ein, s->awr, s->rxn->Q, nuc->e_bins, nuc->n_bins, nuc->mu, s->M,
                                           s->order, distro_int);
            } else {
                ref_fatal("File 4 Reaction Found With Lab Angle Distribution and No Energy Distribution!");
            }
            if (pass == 0)
                for (k = 0; k < n; ++k) result[k] = distro_int[k] * (ONE - f);
            else
                for (k = 0; k < n; ++k) result[k] = result[k] + distro_int[k] * f;
        }
        free(distro_int);
    } else if (s->has_edist) {
        if (s->rxn->scatter_in_cm) {
            unitbase_then(nuc, s, Ein, iE, 1, result);
        } else if (s->has_adist) {
            if (s->law == 9) {
                f = (Ein - A1(s->E_grid, iE)) / (A1(s->E_grid, iE + 1) - A1(s->E_grid, iE));
                distro_int = (double *)malloc(sizeof(double) * n);
                for (pass = 0; pass < 2; ++pass) {
                    for (k = 0; k < n; ++k) distro_int[k] = ZERO;
                    ref_law9_scatter_lab_leg(s->distro[iE - 1 + pass], s->edist_data, Ein, nuc->e_bins, nuc->n_bins,
                                             nuc->mu, s->M, s->order, distro_int);
                    if (pass == 0)
                        for (k = 0; k < n; ++k) result[k] = (ONE - f) * distro_int[k];
                    else
                        for (k = 0; k < n; ++k) result[k] = result[k] + f * distro_int[k];
                }
                free(distro_int);
            } else if (s->law == 4) {
                unitbase_then(nuc, s, Ein, iE, 0, result);
            } else {
                ref_fatal("Associated Edist and Adist, but not law 9");
            }
        } else {
            unitbase_then(nuc, s, Ein, iE, 0, result);
        }
    }
}

/* src/scattdata_header.F90:391-499; distro (order x groups) is fully written */
static void interp_distro(ref_nuclide *nuc, ref_slot *s, double Ein, double *distro)
{
    ref_rxn *rxn = s->rxn;
    int n = s->order * s->groups, k, iE, nuc_iE, n_sig;
    double f, p_valid, sigS;
    const double *sigS_array;

    for (k = 0; k < n; ++k) distro[k] = ZERO;
    if (rxn->MT == REF_ELASTIC) {
        sigS_array = nuc->elastic;
        n_sig = nuc->n_grid;
    } else {
        sigS_array = rxn->sigma;
        n_sig = rxn->n_sigma;
    }

    if (((Ein <= A1(nuc->energy, rxn->threshold)) && (rxn->threshold > 1)) || (Ein > A1(nuc->e_bins, nuc->n_bins))) {
        return;
    } else if (Ein >= A1(nuc->energy, nuc->n_grid)) {
        sigS = A1(sigS_array, n_sig);
        iE = s->NE;
        integrate_distro(nuc, s, Ein, iE - 1, distro);
    } else {
        if (Ein <= A1(nuc->energy, 1))
            nuc_iE = 1;
        else
            nuc_iE = ref_binary_search(nuc->energy, nuc->n_grid, Ein);
        if (A1(nuc->energy, nuc_iE) == A1(nuc->energy, nuc_iE + 1)) nuc_iE = nuc_iE + 1;
        f = (Ein - A1(nuc->energy, nuc_iE)) / (A1(nuc->energy, nuc_iE + 1) - A1(nuc->energy, nuc_iE));
        nuc_iE = nuc_iE - rxn->threshold + 1;
        sigS = (ONE - f) * A1(sigS_array, nuc_iE) + f * A1(sigS_array, nuc_iE + 1);
        if (sigS <= ZERO) return;
        if (Ein < A1(s->E_grid, 1))
            iE = 1;
        else
            iE = ref_binary_search(s->E_grid, s->NE, Ein);
        if (A1(s->E_grid, iE) >= A1(s->E_grid, iE + 1)) iE = iE + 1;
        integrate_distro(nuc, s, Ein, iE, distro);
    }

    if (s->has_edist)
        p_valid = ref_interpolate_tab1(s->p_valid_tab1, Ein);
    else
        p_valid = ONE;
    if (rxn->MT != REF_ELASTIC)
        for (k = 0; k < n; ++k) distro[k] = distro[k] * sigS * p_valid;
}

int ref_nuclide_interp_distro(void *h, int slot, double Ein, double *distro)
{
    ref_nuclide *nuc = (ref_nuclide *)h;
    if (slot < 0 || slot >= nuc->n_slots || !nuc->slots[slot].is_init) return 1;
    interp_distro(nuc, &nuc->slots[slot], Ein, distro);
    return 0;
}

/* chunk of the E_in loops: 100 as src/scatt.F90:631,722; bench.py lowers it for its bounded samples
 * (a few hundred points), which would otherwise fall into one or two chunks */
static int g_omp_chunk = 100;
void ref_set_omp_chunk(int chunk) { g_omp_chunk = chunk > 0 ? chunk : 100; }

static int slot_order(ref_nuclide *nuc)
{
    int i, order = 0;
    for (i = 0; i < nuc->n_slots; ++i)
        if (nuc->slots[i].is_init) order = nuc->slots[i].order; /* inittedSD % order, scatt.F90:143 */
    return order;
}

/* src/scatt.F90:603-675.  el_mat is (order, groups, NE) column-major. */
int ref_nuclide_elastic(void *h, const double *Ein, int NE, double *el_mat, int n_threads)
{
    ref_nuclide *nuc = (ref_nuclide *)h;
    int order = slot_order(nuc), groups = nuc->n_bins - 1, iE;
    size_t col = (size_t)order * groups;
    (void)n_threads;

#pragma omp parallel for schedule(dynamic, g_omp_chunk) num_threads(n_threads > 0 ? n_threads : 1)
    for (iE = 1; iE <= NE; ++iE) {
        int irxn;
        double *out = el_mat + col * (iE - 1);
        if (A1(Ein, iE) <= A1(nuc->e_bins, nuc->n_bins)) {
            memset(out, 0, sizeof(double) * col);
            for (irxn = 0; irxn < nuc->n_slots; ++irxn) {
                ref_slot *s = &nuc->slots[irxn];
                if (!s->is_init) continue;
                if (s->rxn->MT != REF_ELASTIC) continue;
                interp_distro(nuc, s, A1(Ein, iE), out);
            }
        }
    }
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : 1)
    ref_freegas_counters_flush();
    for (iE = 1; iE <= NE; ++iE)
        if (!(A1(Ein, iE) <= A1(nuc->e_bins, nuc->n_bins)) && iE > 1)
            memcpy(el_mat + col * (iE - 1), el_mat + col * (iE - 2), sizeof(double) * col);
    return ref_error_count();
}

/* src/scatt.F90:682-778 */
int ref_nuclide_inelastic(void *h, const double *Ein, int NE, double *inel_mat, double *nuinel_mat, int n_threads)
{
    ref_nuclide *nuc = (ref_nuclide *)h;
    int order = slot_order(nuc), groups = nuc->n_bins - 1, iE;
    size_t col = (size_t)order * groups;
    (void)n_threads;

    memset(inel_mat, 0, sizeof(double) * col * NE);
    if (nuinel_mat) memset(nuinel_mat, 0, sizeof(double) * col * NE);

#pragma omp parallel num_threads(n_threads > 0 ? n_threads : 1)
    {
        double *temp = (double *)malloc(sizeof(double) * col);
#pragma omp for schedule(dynamic, g_omp_chunk)
        for (iE = 1; iE <= NE; ++iE) {
            int irxn;
            size_t k;
            double yield;
            double *out = inel_mat + col * (iE - 1);
            double *nuout = nuinel_mat ? nuinel_mat + col * (iE - 1) : NULL;
            if (A1(Ein, iE) <= A1(nuc->e_bins, nuc->n_bins)) {
                for (irxn = 0; irxn < nuc->n_slots; ++irxn) {
                    ref_slot *s = &nuc->slots[irxn];
                    if (!s->is_init) continue;
                    if (s->rxn->MT == REF_ELASTIC) continue;
                    interp_distro(nuc, s, A1(Ein, iE), temp);
                    for (k = 0; k < col; ++k) out[k] = out[k] + temp[k];
                    if (nuout) {
                        if (s->rxn->yield_tab1)
                            yield = ref_interpolate_tab1(s->rxn->yield_tab1, A1(Ein, iE));
                        else
                            yield = (double)s->rxn->multiplicity;
                        for (k = 0; k < col; ++k) nuout[k] = nuout[k] + yield * temp[k];
                    }
                }
            }
        }
        free(temp);
    }
    for (iE = 1; iE <= NE; ++iE)
        if (!(A1(Ein, iE) <= A1(nuc->e_bins, nuc->n_bins)) && iE > 1) {
            memcpy(inel_mat + col * (iE - 1), inel_mat + col * (iE - 2), sizeof(double) * col);
            if (nuinel_mat) memcpy(nuinel_mat + col * (iE - 1), nuinel_mat + col * (iE - 2), sizeof(double) * col);
        }
    return ref_error_count();
}

/* calc_scatt's preamble of create_Ein_grid (src/scatt.F90:89-140): inel_thresh = the lowest threshold energy of the
 * initialised non-elastic slots (from energy_bins(size)), cutoff = the elastic slot's freegas_cutoff; then egrid_ref.c */
int ref_create_ein_grid(int n_slots, const int *is_init, const int *MT, const double *Q_value, const double *const *E_grid,
                        const int *NE, const double *E_bins, int nb, const double *nuc_grid, int n_grid, double awr,
                        double kT, double cutoff, double thresh, int extend_pts, int inel_extend_pts, double *Ein_el,
                        int *n_el, double *Ein_inel, int *n_inel, int cap);
int ref_nuclide_create_ein_grid(void *h, int extend_pts, int inel_extend_pts, double *Ein_el, int *n_el, double *Ein_inel,
                                int *n_inel, int cap)
{
    ref_nuclide *nuc = (ref_nuclide *)h;
    int ns = nuc->n_slots, i, rc;
    int *is_init = (int *)calloc((size_t)ns + 1, sizeof(int)), *MT = (int *)calloc((size_t)ns + 1, sizeof(int));
    int *NE = (int *)calloc((size_t)ns + 1, sizeof(int));
    double *Q = (double *)calloc((size_t)ns + 1, sizeof(double));
    const double **eg = (const double **)calloc((size_t)ns + 1, sizeof(double *));
    double inel_thresh = A1(nuc->e_bins, nuc->n_bins), cutoff = ZERO;
    for (i = 0; i < ns; ++i) {
        ref_slot *s = &nuc->slots[i];
        is_init[i] = s->is_init;
        if (!s->is_init) continue;
        MT[i] = s->rxn->MT;
        Q[i] = s->rxn->Q;
        NE[i] = s->NE;
        eg[i] = s->E_grid;
        if (s->rxn->MT == REF_ELASTIC)
            cutoff = s->freegas_cutoff;
        else if (A1(nuc->energy, s->rxn->threshold) < inel_thresh)
            inel_thresh = A1(nuc->energy, s->rxn->threshold);
    }
    rc = ref_create_ein_grid(ns, is_init, MT, Q, eg, NE, nuc->e_bins, nuc->n_bins, nuc->energy, nuc->n_grid, nuc->awr, nuc->kT,
                             cutoff, inel_thresh, extend_pts, inel_extend_pts, Ein_el, n_el, Ein_inel, n_inel, cap);
    free(is_init); free(MT); free(NE); free(Q); free(eg);
    return rc;
}

void ref_nuclide_free(void *h)
{
    ref_nuclide *nuc = (ref_nuclide *)h;
    int i, k;
    if (!nuc) return;
    for (i = 0; i < nuc->n_slots; ++i) {
        ref_slot *s = &nuc->slots[i];
        for (k = 0; k < s->NE; ++k) {
            if (s->distro) free(s->distro[k]);
            if (s->Eouts) free(s->Eouts[k]);
            if (s->pdfs) free(s->pdfs[k]);
            if (s->cdfs) free(s->cdfs[k]);
        }
        free(s->distro); free(s->Eouts); free(s->pdfs); free(s->cdfs);
        free(s->NP); free(s->INTT); free(s->E_grid); free(s->edist_data); free(s->p_valid_tab1);
    }
    for (i = 0; i < nuc->n_rxn; ++i) {
        ref_rxn *r = nuc->rxns[i];
        free(r->yield_tab1); free(r->sigma); free(r->ad_energy); free(r->ad_type); free(r->ad_loc); free(r->ad_data);
        free(r);
    }
    free(nuc->rxns); free(nuc->rxn_ids); free(nuc->slots);
    free(nuc->energy); free(nuc->elastic); free(nuc->e_bins); free(nuc->mu);
    free(nuc);
}
