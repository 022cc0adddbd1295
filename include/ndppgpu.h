/*
 * ndppgpu.h -- C-ABI of libndppgpu.so, the B200 (sm_100a) scattering-moment integrator for NDPP.
 *
 * The reference (ndpp/ndpp, Fortran) has no plugin or FFI interface; the boundary is the Fortran
 * procedure seam inside calc_scatt / calc_scattsab (src/scatt.F90:33,543).  Each entry point
 * below replaces one reference routine; the Fortran driver, ndpp.xml handling, ACE parsing,
 * E_in grid construction, tolerance/thinning and output stay in the reference and call these
 * through ISO_C_BINDING (the interface module is shown in INTEGRATION.md).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; ndppgpu_last_error() gives the
 *     text (the Fortran wrapper passes it to fatal_error, src/error.F90:79).  No exceptions
 *     cross the ABI.
 *   - the caller owns every host array; the library copies during the call and keeps nothing.
 *   - index-valued arguments (threshold, locators inside the raw ACE blocks) are 1-based /
 *     ACE-relative exactly as the reference stores them (src/ace_header.F90, src/ace.F90).
 *   - moment arrays are Fortran `mat(L, G, NE)` column-major == C `mat[iE][g][l]`, l contiguous,
 *     L = order+1 for Legendre output (src/scattdata_header.F90:114-118).
 *   - there is no CPU fallback: without a CUDA device every call fails with an error.
 */
#ifndef NDPPGPU_H
#define NDPPGPU_H

#ifdef __cplusplus
extern "C" {
#endif

#define NDPPGPU_ABI_VERSION 3

/* run-time integration parameters: src/global.F90:28-59, defaults src/constants.F90:69-100 */
typedef struct {
    int scatt_type;        /* SCATT_TYPE_LEGENDRE = 0 (TABULAR = 1 is unimplemented in the reference) */
    int order;             /* scatt_order (L-1 for Legendre) */
    int mu_bins;           /* M, points of the uniform mu grid */
    int nuscatter;         /* also build nuinel_mat */
    int ne_per_grp;        /* NE_PER_GRP */
    int adaptive_mu_its;   /* ADAPTIVE_MU_ITS */
    int adaptive_eout_its; /* ADAPTIVE_EOUT_ITS */
    int reserved;
    double sab_threshold;     /* SAB_THRESHOLD */
    double brent_mu_thresh;   /* BRENT_MU_THRESH */
    double adaptive_mu_tol;   /* ADAPTIVE_MU_TOL */
    double adaptive_eout_tol; /* ADAPTIVE_EOUT_TOL */
} ndppgpu_params;

/* counters of the work done since the last reset (ndppgpu_stats) */
typedef struct {
    double kernel_ms;        /* CUDA-event time of all kernels launched by the integrator calls */
    double h2d_bytes;        /* host->device bytes copied by the calls */
    double d2h_bytes;        /* device->host bytes copied by the calls */
    long long launches;      /* kernels launched */
    long long moment_evals;  /* output elements (E_in, g, l) produced */
    long long file4_calls;   /* integrate_file4_cm_leg evaluations (E_in x reaction x table row) */
    long long file6_cm_points; /* inner (group, E_out, mu) points of integrate_file6_cm_leg */
    long long file6_lab_calls;
    long long freegas_tasks; /* adaptive (E_in, row, g, l) free-gas integrations */
    long long sab_columns;
    double file6_cm_ms;      /* CUDA-event time of the dominant kernel (k_file6_cm) alone */
    long long file6_cm_launches;
    double host_call_ms;     /* host wall time spent inside the integrator entry points */
    double host_alloc_ms;    /* ... of which in device allocations (stream-ordered pool) */
    double host_sync_ms;     /* ... of which blocked on the device (count read-backs, result copies) */
    long long freegas_items; /* work items the free-gas sub-integrals were cut into (all generations) */
    long long freegas_kernel_evals; /* free-gas kernel values actually evaluated (shared by the Legendre orders of a group) */
    long long freegas_sab_evals;    /* calc_sab evaluations actually performed (find_FG_mu, once per outgoing energy) */
} ndppgpu_stats_t;

/* ---- context -------------------------------------------------------------------------------- */
/* One context per process and device (the reference's MPI rank / OpenMP team, src/ndpp.F90:934).
 * device < 0 selects the current CUDA device. */
int ndppgpu_init(int device, void **ctx);
int ndppgpu_finalize(void *ctx);
int ndppgpu_last_error(void *ctx /* may be NULL */, char *buf, int len);
int ndppgpu_stats(void *ctx, ndppgpu_stats_t *out, int reset);
int ndppgpu_abi_version(void);
/* the CUDA stream all work of this context is issued on (as cudaStream_t) */
void *ndppgpu_stream(void *ctx);

/* ---- continuous-energy nuclide: replaces the body of calc_scatt between create_Ein_grid and the
 *      output (src/scatt.F90:84-155) ------------------------------------------------------------ */
/* type(Nuclide) scalars + energy grid + elastic xs (src/ace_header.F90:94-112), the group
 * structure and the integration parameters. */
int ndppgpu_nuclide_create(void *ctx, double awr, double kT, double freegas_cutoff, int n_grid,
                           const double *energy, const double *elastic_xs, const double *e_bins,
                           int n_bins /* G+1 */, const ndppgpu_params *params, void **nuc);

/* One call per ScattData slot, in the order calc_scatt fills rxn_data(:) (src/scatt.F90:88-105):
 * each reaction once, then once more per nested energy distribution (edist%next).  Calls that
 * share `rxn_index` describe the same reaction; the reaction-level arguments (MT .. adist_*) are
 * taken from the first of them.  Performs scatt_init (src/scattdata_header.F90:78-271), i.e. the
 * MT / law filter, the isotropic-adist synthesis and the table allocation.
 *   yield_tab1 / p_valid_tab1: TAB1 flattened as [NR, NBT(NR), INT(NR), NP, x(NP), y(NP)]
 *                              (src/interpolation.F90:24-60); NULL,0 when absent.
 *   adist_*:  DistAngle (src/ace_header.F90:14-24); edist_data: DistEnergy%data (raw DLW block). */
int ndppgpu_nuclide_add_reaction(void *nuc, int rxn_index, int MT, double Q_value, int threshold,
                                 int scatter_in_cm, int has_angle_dist, int has_energy_dist, int law,
                                 int multiplicity, const double *yield_tab1, int n_yield,
                                 const double *sigma, int n_sigma, const double *p_valid_tab1,
                                 int n_pvalid, const double *adist_energy, const int *adist_type,
                                 const int *adist_loc, int n_adist_e, const double *adist_data,
                                 int n_adist_data, const double *edist_data, int n_edist_data);

/* scatt_convert_distro for every slot (src/scattdata_header.F90:325-382): ACE -> uniform-mu tables,
 * computed on the device. */
int ndppgpu_convert_distro(void *nuc);

/* calc_elastic_grid (src/scatt.F90:603-675): el_mat[NE][G][L]. */
int ndppgpu_elastic(void *nuc, const double *Ein, int NE, double *el_mat);
/* calc_inelastic_grid (src/scatt.F90:682-778): inel_mat[NE][G][L]; nuinel_mat may be NULL. */
int ndppgpu_inelastic(void *nuc, const double *Ein, int NE, double *inel_mat, double *nuinel_mat);
/* Same, with E_in and the result left in device memory (pointers are device pointers of this
 * context's device).  Used to shard E_in ranges over GPUs and gather with NCCL. */
int ndppgpu_elastic_dev(void *nuc, const double *d_Ein, int NE, double *d_el_mat);
int ndppgpu_inelastic_dev(void *nuc, const double *d_Ein, int NE, double *d_inel_mat, double *d_nuinel_mat);

/* introspection used by the parity tests */
int ndppgpu_nuclide_n_slots(void *nuc);
/* info[8] = is_init, NE, law, has_adist, has_edist, order(L), groups, MT */
int ndppgpu_nuclide_slot_info(void *nuc, int slot, int *info);
int ndppgpu_nuclide_slot_row_np(void *nuc, int slot, int iE /* 1-based */);
/* row iE of the converted tables: distro[NP][M] (== Fortran data(M,NP)), Eouts/pdf/cdf[NP] */
int ndppgpu_nuclide_get_table(void *nuc, int slot, int iE, double *distro, double *Eouts, double *pdf,
                              double *cdf, int *INTT);
/* Overwrite row iE of a slot's converted table with host values distro[NP][M].  Lets the Fortran side
 * keep convert_distro on the host (its tables then are bit-identical to the reference's own libm),
 * and lets the parity tests separate the integrators from the table conversion. */
int ndppgpu_nuclide_set_table(void *nuc, int slot, int iE, const double *distro);
/* ndppgpu_elastic + ndppgpu_inelastic as calc_scatt calls them one after the other (src/scatt.F90:143-150), in one call:
 * the elastic matrices are copied to the host while the inelastic kernels run.  Either grid may be empty (NE 0, null
 * pointers); nuinel_mat may be null.  Same results as the two calls. */
int ndppgpu_calc_scatt(void *nuc, const double *Ein_el, int NE_el, double *el_mat, const double *Ein_inel, int NE_inel,
                       double *inel_mat, double *nuinel_mat);
/* ScattData%interp_distro (src/scattdata_header.F90:391-499) of one slot at NE incoming energies:
 * distro[NE][G][L], scaled by sigma_s * p_valid for non-elastic reactions exactly as the reference. */
int ndppgpu_interp_distro(void *nuc, int slot, const double *Ein, int NE, double *distro);
/* create_Ein_grid (src/scatt.F90:166-236: combine_Eins :246, add_elastic_Eins :311, add_one_more_point :426,
 * add_inelastic_Eins :456) with calc_scatt's inel_thresh and cutoff (:89-120), on the device, after
 * ndppgpu_convert_distro as in the reference (:107-139).  extend_pts / inel_extend_pts = EXTEND_PTS / INEL_EXTEND_PTS
 * (src/constants.F90:85-86: 50 / 30).  The grids stay on the device; *n_el / *n_inel receive their lengths (*n_inel = 0
 * for a nuclide with elastic scattering only: Ein_inel stays unallocated in the reference).  The points the reference
 * places with log / exp carry the bits of the host's C library (csrc/libm_exact.cuh).  *status (nullable): bit 1 a
 * critical energy of add_inelastic_Eins was NaN and its points were left out (as the reference's merge ends up doing:
 * it drops an all-NaN array), bit 2 an input array held a repeated value (dropped here; the reference's merge keeps
 * some of them). */
int ndppgpu_nuclide_create_ein_grid(void *nuc, int extend_pts, int inel_extend_pts, int *n_el, int *n_inel, int *status);
/* which = 0: Ein_el, 1: Ein_inel.  Ein (nullable): host array of the length returned above; d_Ein (nullable) receives
 * the device pointer (valid until the next create_ein_grid / nuclide_free) for ndppgpu_elastic_dev / _inelastic_dev. */
int ndppgpu_nuclide_ein_grid(void *nuc, int which, double *Ein, const double **d_Ein);
/* ScattData%clear for all slots (src/scatt.F90:153-155) */
int ndppgpu_nuclide_free(void *nuc);

/* ---- thermal S(a,b): replaces integrate_sab_el/_inel + combine_sab_grid in calc_scattsab
 *      (src/scatt.F90:573-591; src/sab.F90:21-454) ---------------------------------------------- */
/* type(SAlphaBeta) (src/ace_header.F90:201-235), flattened with the Fortran column-major order:
 *   inelastic_e_out(NEo,NEi) -> [iEin][iEout];  inelastic_mu(n_mu,NEo,NEi) -> [iEin][iEout][imu]
 *   continuous mode (secondary_mode = 2): cont_n_e_out[NEi] and the rows of e_out / e_out_pdf /
 *   mu(n_mu, NEo_i) concatenated;  elastic_mu(n_mu,NEe) -> [iEin][imu]. */
int ndppgpu_sab_create(void *ctx, double awr, double kT, double threshold_inelastic, double threshold_elastic,
                       int n_inelastic_e_in, int n_inelastic_e_out, int n_inelastic_mu, int secondary_mode,
                       const double *inelastic_e_in, const double *inelastic_sigma, const double *inelastic_e_out,
                       const double *inelastic_mu, const int *cont_n_e_out, const double *cont_e_out,
                       const double *cont_pdf, const double *cont_mu, int elastic_mode, int n_elastic_e_in,
                       int n_elastic_mu, const double *elastic_e_in, const double *elastic_P,
                       const double *elastic_mu, void **sab);
/* calc_scattsab minus sab_egrid: scatt_mat[NE][G][order+1].  el_out / inel_out (nullable) receive
 * the partial integrals of integrate_sab_el / integrate_sab_inel. */
int ndppgpu_sab(void *sab, const double *e_bins, int n_bins, int scatt_type, int order, const double *Ein, int NE,
                double *scatt_mat, double *el_out, double *inel_out);
int ndppgpu_sab_dev(void *sab, const double *e_bins, int n_bins, int scatt_type, int order, const double *d_Ein,
                    int NE, double *d_scatt_mat);
/* sab_egrid (src/sab.F90:460-568) on the device: merged incoming grids + group structure + the incoming energies at
 * which an outgoing energy crosses a group edge, cut at the tables' top, extend_pts (EXTEND_PTS) points inside every
 * interval unless sab_epts_per_bin (SAB_EPTS_PER_BIN, src/global.F90) is 0.  *n = length; ndppgpu_sab_ein_grid as above. */
int ndppgpu_sab_egrid(void *sab, const double *e_bins, int n_bins, int sab_epts_per_bin, int extend_pts, int *n, int *status);
int ndppgpu_sab_ein_grid(void *sab, double *Ein, const double **d_Ein);
int ndppgpu_sab_free(void *sab);

/* ---- leaf check of the device Legendre helpers: integrals[n][L] = calc_int_pn_tablelin(L, xlow, xhigh,
 *      flow, fhigh) (src/legendre.F90:22) and pn[n][L] = calc_pn(l, xlow) (:349) for n inputs ------- */
/* ---- the two steps that follow the integrator in the reference's driver (src/ndpp.F90:611-648) ----
 *
 * ndppgpu_apply_tol  replaces apply_tol_scatt(data, tol) (src/scatt.F90:786-818): groups whose P0 lies in
 *                    (0, tol) are zeroed for every order and the column is renormalised to its original
 *                    sum of P0.  mat is [NE][G][L] (Fortran data(L, G, NE)), modified in place.
 * ndppgpu_thin_grid  replaces thin_grid(xout, yout, tokeep, tol, compression, maxerr [, yout2])
 *                    (src/thin.F90:19-47): greedy thinning of the E_in grid with log-x interpolation; x, y1
 *                    and (if not NULL) y2 are compacted in place to the *n_kept points kept, as the Fortran
 *                    re-allocates its arrays.  compression is the reference's; max_abs_err is the plain
 *                    maximum of |interpolated - y| over the accepted tests (the reference's maxerr mixes a
 *                    relative comparison with an absolute store, src/thin.F90:127-131, and is not reproduced).
 * The *_dev forms take device pointers (d_keep: NE ints of scratch receiving the kept indices) so that the
 * tolerance and the thinning can run before the moment arrays are gathered or copied to the host. */
int ndppgpu_apply_tol(void *ctx, double *mat, int NE, int G, int L, double tol);
int ndppgpu_apply_tol_dev(void *ctx, double *d_mat, int NE, int G, int L, double tol);
int ndppgpu_thin_grid(void *ctx, double *x, double *y1, double *y2, int NE, int GL, const double *tokeep, int n_tokeep,
                      double tol, int *n_kept, double *compression, double *max_abs_err);
int ndppgpu_thin_grid_dev(void *ctx, const double *d_x, const double *d_y1, const double *d_y2, int NE, int GL,
                          const double *tokeep, int n_tokeep, double tol, int *d_keep, int *n_kept, double *compression,
                          double *max_abs_err);
int ndppgpu_gather_columns_dev(void *ctx, const double *d_src, const int *d_keep, int n_kept, int width, double *d_dst);

/* ---- fission-spectrum (chi) integration: replaces the E_in loop of calc_chi (src/chi.F90:120-153), i.e.
 *      ChiData % integrate / prob / beta (src/chidata_header.F90:143-494) with nu_total / nu_delayed
 *      (src/fission.F90:18-103), for one nuclide on the merged incoming-energy grid the caller built
 *      (src/chi.F90:96-112).  A slot is one ChiData object: the prompt laws first, in the reference's order
 *      (fission reaction by fission reaction, nested laws in chain order), then one delayed law per precursor
 *      group.  energy / fission are nuc % energy and nuc % fission; nu_*_type are NU_NONE 0, NU_POLYNOMIAL 1,
 *      NU_TABULAR 2 with nu_*_data as the reference stores them; precursor_data is nuc % nu_d_precursor_data.
 *      Outputs in Fortran order: chi_total(G, NE), chi_prompt(G, NE), chi_delay(G, NE, n_precursor). ------- */
typedef struct {
    int law;        /* edist % law */
    int delayed;    /* 0 = prompt, 1 = delayed */
    int precursor;  /* 1-based precursor group (delayed) */
    int threshold;  /* rxn % threshold, 1-based (prompt) */
    int use_pvalid; /* associated(edist % next) .and. edist % p_valid % n_regions > 0 */
    int n_sigma;    /* size of the cross section chi_prob reads (nuc % fission for MT 18, rxn % sigma otherwise) */
    int sigma_off;  /* offsets, in doubles, into `pool`: that cross section, edist % data, p_valid as a TAB1 */
    int data_off;
    int pvalid_off;
    int reserved;
} ndppgpu_chi_slot;
int ndppgpu_chi(void *ctx, int n_grid, const double *energy, const double *fission, int nu_t_type,
                const double *nu_t_data, int n_nu_t, int nu_d_type, const double *nu_d_data, int n_nu_d,
                int n_precursor, const double *precursor_data, int n_precursor_data, int n_slots,
                const ndppgpu_chi_slot *slots, const double *pool, int n_pool, const double *e_bins, int n_bins,
                const double *Ein, int NE, double *chi_total, double *chi_prompt, double *chi_delay);

/* calc_elastic_grid / calc_inelastic_grid followed by apply_tol_scatt and (if thin_tol > 0) thin_grid, i.e.
 * src/ndpp.F90:607-648 for one matrix set, with only the kept columns copied back.  Ein is in/out: NE points
 * in, the *n_kept points kept out; the matrices must have room for NE columns and hold n_kept on return. */
int ndppgpu_elastic_thinned(void *nuc, double *Ein, int NE, double print_tol, double thin_tol, const double *tokeep,
                            int n_tokeep, double *el_mat, int *n_kept, double *compression, double *max_abs_err);
int ndppgpu_inelastic_thinned(void *nuc, double *Ein, int NE, double print_tol, double thin_tol, const double *tokeep,
                              int n_tokeep, double *inel_mat, double *nuinel_mat /* nullable */, int *n_kept,
                              double *compression, double *max_abs_err);

/* Self-test of the shared-reciprocal division (csrc/legendre.cuh: FastDiv) against the IEEE operator on
 * per_thread random operand pairs per GPU thread.  counts2 = {pairs, mismatches}; mismatches must be zero. */
int ndppgpu_test_exact_math(void *ctx, unsigned long long seed, int per_thread, unsigned long long *counts2);

/* The device's sinh / cosh / expm1 / exp (csrc/libm_exact.cuh: a restatement of the algorithms of the host C library
 * the reference's Fortran calls, GNU libc 2.39 x86-64 FMA builds) at n host arguments; fn = 0 exp, 1 expm1, 2 sinh,
 * 3 cosh.  convert_file6's Law 44 (src/scattdata_header.F90:822-831) is built from these, and the parity tests compare
 * them bit for bit with the running libm.  fn 10 .. 12: the branch-free primitives of the free-gas kernel (exp for
 * -708 < x <= 0, sqrt, x[i] / x[i ^ 1]; n even for 12): where a primitive's range guard fails the library operation is
 * returned, so comparing y with exp / sqrt / division tests exactly the guarded results. */
int ndppgpu_eval_libm(void *ctx, int fn, const double *x, long long n, double *y);

int ndppgpu_test_legendre(void *ctx, int n, int L, const double *xlow, const double *xhigh, const double *flow,
                          const double *fhigh, double *integrals, double *pn);

/* ==== several GPUs =====================================================================================
 * Replaces the reference's own distribution of a library: partition_work (static contiguous blocks of nuclides per
 * MPI rank, src/ndpp.F90:934-950), the nuclide loop (:549) and the hand-back of the results (:839-864).
 *
 * A *group* is the set of GPUs that work together: every GPU of this process (ndppgpu_group_init: one host thread per
 * device inside the library, ncclCommInitAll) or one GPU per process (ndppgpu_group_init_rank: ncclCommInitRank with
 * an id made by ndppgpu_group_unique_id on rank 0 and broadcast by the caller, e.g. MPI_Bcast).  Global rank 0 is the
 * root: it receives the assembled matrices.  In the one-GPU-per-process form every ndppgpu_group_* / ndppgpu_library_*
 * call is collective (all ranks, same arguments); result arrays may be NULL on the other ranks.
 * NCCL is loaded on demand (libnccl.so.2, or $NDPPGPU_NCCL_LIB); a group of one device needs none. */
int ndppgpu_group_init(int n_devices /* <= 0: all */, const int *devices /* NULL: 0..n-1 */, void **group);
int ndppgpu_group_unique_id(void *id128 /* 128 bytes out */);
int ndppgpu_group_init_rank(int device, int rank, int world, const void *id128, void **group);
int ndppgpu_group_info(void *group, int *world, int *n_local, int *first_rank);
/* borrowed context of a local device (ndppgpu_stats, ndppgpu_stream); owned by the group */
void *ndppgpu_group_ctx(void *group, int local_index);
/* bytes the root received over NCCL since the last reset */
long long ndppgpu_group_gathered_bytes(void *group, int reset);
int ndppgpu_group_finalize(void *group);

/* A nuclide replicated on every device of the group; same arguments as ndppgpu_nuclide_create / _add_reaction /
 * ndppgpu_convert_distro / ndppgpu_elastic / ndppgpu_inelastic.  The E_in grid is dealt cyclically over the devices
 * (column i to device i mod N), every device integrates its columns, the columns are gathered to the root device with
 * ncclSend / ncclRecv over NVLink, put back in order, and the top-of-grid rule (src/scatt.F90:669,770) is applied on
 * the root.  The matrices equal the one-device call's bit for bit. */
int ndppgpu_group_nuclide_create(void *group, double awr, double kT, double freegas_cutoff, int n_grid,
                                 const double *energy, const double *elastic_xs, const double *e_bins, int n_bins,
                                 const ndppgpu_params *params, void **gnuc);
int ndppgpu_group_nuclide_add_reaction(void *gnuc, int rxn_index, int MT, double Q_value, int threshold,
                                       int scatter_in_cm, int has_angle_dist, int has_energy_dist, int law,
                                       int multiplicity, const double *yield_tab1, int n_yield, const double *sigma,
                                       int n_sigma, const double *p_valid_tab1, int n_pvalid,
                                       const double *adist_energy, const int *adist_type, const int *adist_loc,
                                       int n_adist_e, const double *adist_data, int n_adist_data,
                                       const double *edist_data, int n_edist_data);
int ndppgpu_group_convert_distro(void *gnuc);
int ndppgpu_group_elastic(void *gnuc, const double *Ein, int NE, double *el_mat /* root */);
int ndppgpu_group_inelastic(void *gnuc, const double *Ein, int NE, double *inel_mat, double *nuinel_mat /* root */);
int ndppgpu_group_nuclide_free(void *gnuc);
/* The same in pieces, for callers that keep grids and results on the devices: set_grids deals and uploads the grids once
 * (a negative count leaves that grid as it is); integrate (what = 1 elastic, 2 inelastic, 3 both) only
 * enqueues -- kernels on every device's stream, the gather and the assembly on side streams, two result buffers in
 * turn, so the gather of one call overlaps the kernels of the next; sync waits; fetch copies the latest assembled
 * matrices to host arrays on the root; result_dev is their device address on the root device (0 elastic, 1 inelastic,
 * 2 nu-inelastic), valid until the second next integrate. */
int ndppgpu_group_set_grids(void *gnuc, const double *Ein_el, int NE_el, const double *Ein_inel, int NE_inel);
int ndppgpu_group_integrate(void *gnuc, int what);
int ndppgpu_group_sync(void *gnuc);
/* device-side join: each device's stream (ndppgpu_stream of its context) waits for the gathers enqueued so far */
int ndppgpu_group_join(void *gnuc);
int ndppgpu_group_fetch(void *gnuc, double *el_mat, double *inel_mat, double *nuinel_mat);
void *ndppgpu_group_result_dev(void *gnuc, int matrix);

/* ---- a whole library: work items (nuclide, matrix, E_in tile), cost-weighted, longest-processing-time-first ------------
 * ndppgpu_plan_library is host code (no GPU needed).  A shape is what the cost model needs to know of a nuclide; the
 * cost of a tile is the algorithmic-flop count of SURVEY 8d (F_A per open level, F_B for the CM continuum, the counted
 * F_E per free-gas column) with E_in taken log-uniform between e_lo and e_hi -- it steers the balance only, never a
 * result.  policy 0: LPT with setup_cost (model flops) charged per (device, nuclide) opened; policy 1: the reference's
 * static contiguous blocks of nuclides.  Items come back sorted by (rank, nuclide, matrix, tile). */
typedef struct {
    int index;                      /* caller's nuclide number */
    int n_el, n_inel;               /* E_in points of the two grids */
    int n_levels;                   /* discrete levels (file-4 CM integrations) */
    int has_cont;                   /* Law 44 / 61 continuum in the CM frame (file-6 CM) */
    int freegas_points;             /* elastic E_in below the free-gas cutoff */
    double cont_threshold, e_lo, e_hi;
    const double *level_thresholds; /* [n_levels] MeV */
} ndppgpu_shape;
typedef struct {
    int nuclide, matrix /* 0 elastic, 1 inelastic */, tile, n_tiles, rank;
    int rows;                       /* E_in points of the tile: ndppgpu_tile_bounds(NE, tile, n_tiles) (set by the caller) */
    double cost;
} ndppgpu_item;
int ndppgpu_plan_library(const ndppgpu_shape *shapes, int n_shapes, int G, int L, int M, int K, int tile_rows, int world,
                         double setup_cost, int policy, ndppgpu_item *items_out, int max_items, int *n_items,
                         double *imbalance /* max device load / mean */);
void ndppgpu_tile_bounds(int n, int tile, int n_tiles, int *lo, int *hi);

/* Runs a plan.  Every device walks its items; when it meets a new nuclide it calls open(user, nuclide, ctx, ...) -- the
 * caller parses / builds that nuclide on the given context (ndppgpu_nuclide_create + _add_reaction; convert_distro is
 * called by the library if the caller has not) and returns the handle and its two E_in grids -- integrates the tiles in
 * place into one result buffer per device, and calls close (NULL: ndppgpu_nuclide_free).  open / close are called from
 * the device's worker thread.  Then one NCCL gather brings every buffer to the root device, where ndppgpu_library_fetch
 * assembles a matrix (0 elastic, 1 inelastic, 2 nu-inelastic) of a nuclide into a host array. */
typedef int (*ndppgpu_open_fn)(void *user, int nuclide, void *ctx, void **nuc, const double **Ein_el, int *NE_el,
                               const double **Ein_inel, int *NE_inel);
typedef int (*ndppgpu_close_fn)(void *user, int nuclide, void *nuc);
typedef struct {
    double wall_s, compute_s, gather_s;      /* host wall time: all, until the slowest device finished, the gather */
    double open_s_max, integrate_s_max;      /* slowest local device: opening nuclides / integrating tiles */
    double kernel_s_max, kernel_s_sum;       /* CUDA-event kernel time: slowest local device, sum over local devices */
    long long moment_evals;                  /* whole plan */
    int items, opens;                        /* items of the plan; nuclides opened by the local devices */
    double device_s_max;                     /* CUDA-event time from a device's first upload to the end of its part of the
                                                gather, slowest local device */
    double alloc_s_max;                      /* host time a local device spent allocating its result buffer */
    double reserved[2];
} ndppgpu_library_report;
int ndppgpu_library_create(void *group, int G, int L, int nuscatter, const ndppgpu_item *items, int n_items, void **lib);
int ndppgpu_library_run(void *lib, ndppgpu_open_fn open, ndppgpu_close_fn close, void *user, ndppgpu_library_report *rep);
int ndppgpu_library_fetch(void *lib, int nuclide, int matrix, const double *Ein, int NE, double e_top, double *mat);
int ndppgpu_library_report_get(void *lib, ndppgpu_library_report *rep);
int ndppgpu_library_free(void *lib);

/* ---- device micro-benchmark: sustained FP64 FMA rate of this GPU, used as the roofline
 *      denominator (MEASURED_PEAKS.json holds no FP64 figure) ----------------------------------- */
int ndppgpu_measure_fp64_peak(void *ctx, double seconds, double *tflops);

#ifdef __cplusplus
}
#endif
#endif
