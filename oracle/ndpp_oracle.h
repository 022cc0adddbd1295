/*
 * ORACLE (test infrastructure only).  CPU restatement, in plain C, of the reference's
 * scattering-moment integrator (src/scattdata_header.F90, src/scatt.F90:603-778,
 * src/freegas.F90, src/sab.F90, src/legendre.F90, src/search.F90, src/interpolation.F90,
 * src/array_merge.F90).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product (ndpp_b200/, include/) never does.
 *
 * Parity status: the reference is Fortran and no Fortran compiler exists in the build
 * container, so the oracle cannot be checked against the reference binary.  It is pinned by
 * the reference's own known-answer tests (tests/test_scatt/test_scattdata.F90 and the Sage
 * worksheets beside it) -- see tests/test_oracle_golden.py.  Routines for which the reference
 * holds no test (free gas, S(a,b), unit-base, file6_cm_leg, file6_lab_leg, law 9) are "parity
 * unpinned"; they are held by independent evaluations instead (tests/test_oracle_golden.py): the
 * closed-form free-gas kernel for A = 1 and general A and a double quadrature of the free-gas law for
 * P1..P3, numpy evaluations of every S(a,b) routine, the heavy-target limit of unit-base +
 * file6_cm_leg on separable tables, numerical quadrature of the law-9 spectrum, and a hand evaluation
 * of the Fortran text of file6_lab_leg.
 *
 * Index convention: all table indices handed around inside the oracle are 1-based, exactly as
 * in the Fortran text; arrays are read through the A1() accessor.
 */
#ifndef NDPP_ORACLE_H
#define NDPP_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* src/constants.F90 */
#define REF_PI 3.1415926535898      /* :35  (truncated on purpose) */
#define REF_FP_PRECISION 1e-14      /* :22 */
#define REF_INFINITY 1.7976931348623157e308 /* huge(0.0_8), :36 */
#define REF_MIN_EIN 1e-14           /* :109 */
#define REF_NUM_EP 32               /* :116 */
#define REF_R_NUM_EP (1.0 / 32.0)   /* :118 */

enum { REF_HISTOGRAM = 1, REF_LINEAR_LINEAR = 2, REF_LINEAR_LOG = 3, REF_LOG_LINEAR = 4, REF_LOG_LOG = 5 };
enum { REF_ANGLE_ISOTROPIC = 1, REF_ANGLE_32_EQUI = 2, REF_ANGLE_TABULAR = 3 };
enum { REF_SCATT_TYPE_LEGENDRE = 0, REF_SCATT_TYPE_TABULAR = 1 };
enum { REF_SAB_SECONDARY_EQUAL = 0, REF_SAB_SECONDARY_SKEWED = 1, REF_SAB_SECONDARY_CONT = 2 };
enum { REF_SAB_ELASTIC_DISCRETE = 3, REF_SAB_ELASTIC_EXACT = 4 };
#define REF_ELASTIC 2

#define A1(arr, i) ((arr)[(i)-1])

/* run-time integration parameters, src/global.F90:28-59 */
typedef struct {
    int scatt_type;        /* 0 = Legendre */
    int order;             /* scatt_order (L-1 for Legendre) */
    int mu_bins;           /* M */
    int nuscatter;         /* build nuinel_mat */
    int ne_per_grp;        /* NE_PER_GRP */
    int adaptive_mu_its;   /* ADAPTIVE_MU_ITS */
    int adaptive_eout_its; /* ADAPTIVE_EOUT_ITS */
    int reserved;
    double sab_threshold;     /* SAB_THRESHOLD */
    double brent_mu_thresh;   /* BRENT_MU_THRESH */
    double adaptive_mu_tol;   /* ADAPTIVE_MU_TOL */
    double adaptive_eout_tol; /* ADAPTIVE_EOUT_TOL */
} ref_params;

/* ---- error state (fatal_error in the reference aborts; here it latches) ---- */
int ref_error_count(void);
const char *ref_error_message(void);
void ref_error_clear(void);

/* ---- leaf routines (exported for the known-answer tests) ---- */
double ref_calc_pn(int n, double x);
void ref_calc_int_pn_tablelin(int n, double xlow, double xhigh, double flow, double fhigh, double *integrals);
int ref_binary_search(const double *array, int n, double val);
double ref_interpolate_tab1(const double *data, double x);
int ref_merge(const double *a, int na, const double *b, int nb, double *result);
double ref_tolab(double R, double w);
void ref_convert_file4(int iE, const double *mu, int M, const double *ad_energy, const int *ad_type,
                       const int *ad_location, const double *ad_data, double *distro);
int ref_convert_file6(int iE, const double *mu, int M, int law, const double *data, int *INTT, int *NP_out,
                      double *Eouts, double *pdf, double *cdf, double *distro);
void ref_integrate_file4_cm_leg(const double *fw, double Ein, double awr, double Q, const double *E_bins, int nbins,
                                const double *w, int M, int order, double *distro);
void ref_integrate_file6_cm_leg(const double *fEmu, const double *mu, int M, double Ein, double awr,
                                const double *Eout, int NEout, int INTT, const double *thispdf,
                                const double *E_bins, int nbins, int order, int ne_per_grp, double *distro);
void ref_integrate_file6_lab_leg(const double *fEmu, const double *mu, int M, const double *Eout, int NEout, int INTT,
                                 const double *thispdf, const double *E_bins, int nbins, int order, double *distro);
void ref_law9_scatter_lab_leg(const double *fmu, const double *edist_data, double Ein, const double *E_bins,
                              int nbins, const double *mu, int M, int order, double *distro);
int ref_cast_to_unitbase(const double *Eout, int n, double *ub_grid);
int ref_interp_unitbase(double Ein, const double *ub1, int nub1, const double *Eout1, int n1, const double *pdf1,
                        int INTT1, const double *fEmu1, double Ei1, const double *ub2, int nub2, const double *Eout2,
                        int n2, const double *pdf2, int INTT2, const double *fEmu2, double Ei2, int M, double *Eout,
                        double *pdf, int *INTT, double *fEmu);

/* free gas, src/freegas.F90 */
void ref_integrate_freegas_leg(double Ein, double A, double kT, const double *fEmu, const double *mu, int M,
                               const double *E_bins, int nbins, int order, const ref_params *p, double *distro);
double ref_calc_sab(double A, double kT, double Ein, double Eout, double beta, double mu);
double ref_calc_fgk(double awr, double kT, double Ein, double Eout, int l, double mu, const double *fEmu,
                    const double *global_mu, int M);
void ref_find_FG_mu(double A, double kT, double Ein, double Eout, const ref_params *p, double *mu2);
double ref_brent_mu(double awr, double kT, double Ein, double Eout, double beta, double thresh, double lo, double hi,
                    const ref_params *p);
void ref_calc_FG_Eout_bounds(double A, double kT, double Ein, double *Eout_lo, double *Eout_hi);
/* counters for algorithmic-flop accounting (F_E) */
void ref_freegas_counters(long long *n_fgk, long long *n_sab, int reset);
void ref_freegas_counters_flush(void);
void ref_fatal(const char *msg);

/* ---- nuclide level (mirrors calc_scatt's use of ScattData) ---- */
void *ref_nuclide_create(double awr, double kT, double freegas_cutoff, int n_grid, const double *energy,
                         const double *elastic_xs, const double *e_bins, int n_bins, const ref_params *p);
int ref_nuclide_add_reaction(void *nuc, int rxn_index, int MT, double Q, int threshold, int scatter_in_cm,
                             int has_angle_dist, int has_energy_dist, int law, int multiplicity,
                             const double *yield_tab1, int n_yield, const double *sigma, int n_sigma,
                             const double *p_valid_tab1, int n_pvalid, const double *adist_energy,
                             const int *adist_type, const int *adist_loc, int n_adist_e, const double *adist_data,
                             int n_adist_data, const double *edist_data, int n_edist_data);
int ref_nuclide_n_slots(void *nuc);
int ref_nuclide_slot_info(void *nuc, int slot, int *info /* [8]: is_init, NE, law, has_adist, has_edist, order, groups, MT */);
int ref_nuclide_slot_row_np(void *nuc, int slot, int iE);
int ref_nuclide_slot_egrid(void *nuc, int slot, double *E_grid);
int ref_nuclide_convert_distro(void *nuc);
/* table access: row iE (1-based) of slot; distro is M x NP column-major */
int ref_nuclide_get_table(void *nuc, int slot, int iE, double *distro, double *Eouts, double *pdf, double *cdf,
                          int *INTT);
int ref_nuclide_set_table(void *nuc, int slot, int iE, const double *distro);
int ref_nuclide_interp_distro(void *nuc, int slot, double Ein, double *distro);
int ref_nuclide_elastic(void *nuc, const double *Ein, int NE, double *el_mat, int n_threads);
int ref_nuclide_inelastic(void *nuc, const double *Ein, int NE, double *inel_mat, double *nuinel_mat, int n_threads);
void ref_nuclide_free(void *nuc);
void ref_set_omp_chunk(int chunk);

/* ---- S(a,b), src/sab.F90 + calc_scattsab (src/scatt.F90:543-596) ---- */
void *ref_sab_create(double awr, double kT, double threshold_inelastic, double threshold_elastic, int n_inelastic_e_in,
                     int n_inelastic_e_out, int n_inelastic_mu, int secondary_mode, const double *inelastic_e_in,
                     const double *inelastic_sigma, const double *inelastic_e_out, const double *inelastic_mu,
                     const int *cont_n_e_out, const double *cont_e_out, const double *cont_pdf, const double *cont_mu,
                     int elastic_mode, int n_elastic_e_in, int n_elastic_mu, const double *elastic_e_in,
                     const double *elastic_P, const double *elastic_mu);
int ref_sab_calc(void *sab, const double *e_bins, int n_bins, int order, const double *Ein, int NE, double *scatt_mat,
                 double *el_out, double *inel_out, int n_threads);
/* TABULAR output (project-defined, parity unpinned): `order` equal-width cosine bins */
int ref_sab_calc_tabular(void *sab, const double *e_bins, int n_bins, int order, const double *Ein, int NE,
                         double *scatt_mat, double *el_out, double *inel_out, int n_threads);
void ref_sab_free(void *sab);

/* ---- the two steps after the integrator: apply_tol_scatt (src/scatt.F90:786-818), thin_grid (src/thin.F90) ---- */
void ref_apply_tol_scatt(double *data, int L, int G, int NE, double tol);
int ref_thin_grid(const double *x, const double *y1, const double *y2, int NE, int GL, const double *tokeep, int n_tokeep,
                  double tol, int *keep, double *compression, double *maxerr, double *max_abs);

/* ---- chi (fission spectrum) integration: calc_chi (src/chi.F90:21-163), src/chidata_header.F90:143-494 ---- */
typedef struct {
    int law;        /* edist % law */
    int delayed;    /* 0 = prompt law of a fission reaction, 1 = delayed-neutron law of a precursor group */
    int precursor;  /* 1-based precursor group (delayed) */
    int threshold;  /* rxn % threshold, 1-based (prompt) */
    int use_pvalid; /* associated(edist % next) .and. p_valid % n_regions > 0 */
    int n_sigma;    /* length of the cross section the probability is taken from */
    int sigma_off;  /* offsets (in doubles) into the pool: sigma, edist % data, flattened p_valid TAB1 */
    int data_off;
    int pvalid_off;
    int reserved;
} ref_chi_slot;
int ref_calc_chi(int n_grid, const double *energy, const double *fission, int nu_t_type, const double *nu_t_data,
                 int nu_d_type, const double *nu_d_data, int n_precursor, const double *precursor_data, int n_slots,
                 const ref_chi_slot *slots, const double *pool, const double *E_bins, int n_bins, const double *Ein_grid,
                 int NE, double *chi_total, double *chi_prompt, double *chi_delay);

#ifdef __cplusplus
}
#endif
#endif
