!===============================================================================
! NDPP_GPU -- ISO_C_BINDING interface to libndppgpu.so (include/ndppgpu.h), the
! B200 implementation of NDPP's scattering-moment integrator.
!
! Drop this file into the reference's src/ directory, add ndpp_gpu.o to the
! objects of src/Makefile (it depends on error.o, ace_header.o, endf_header.o)
! and link with
!     -L<repo>/ndpp_b200/csrc -lndppgpu -Wl,-rpath,<repo>/ndpp_b200/csrc
! INTEGRATION.md shows the three procedure bodies of src/scatt.F90 that call it.
!
! NOTE: the image this project is built in has no Fortran compiler (gfortran,
! flang, nvfortran, ifort, ifx, lfortran, f2c probed: all absent), so this
! module has been written against the reference sources by inspection and has
! not been compiled; the identical call sequence is exercised by the C++ twin
! (include/ndpp_host.hpp, tools/ndpp_calc_scatt.cpp) and the Python twin
! (ndpp_b200/scatt.py), which the parity tests drive.
!===============================================================================

module ndpp_gpu

  use, intrinsic :: iso_c_binding

  implicit none
  private

  public :: ndppgpu_params, ndppgpu_chi_slot, gpu_ctx, gpu_start, gpu_stop, gpu_check, tab1_flat
  public :: ndppgpu_init, ndppgpu_finalize, ndppgpu_last_error
  public :: ndppgpu_nuclide_create, ndppgpu_nuclide_add_reaction, ndppgpu_convert_distro
  public :: ndppgpu_elastic, ndppgpu_inelastic, ndppgpu_nuclide_free
  public :: ndppgpu_calc_scatt
  public :: ndppgpu_nuclide_create_ein_grid, ndppgpu_nuclide_ein_grid, ndppgpu_sab_egrid, ndppgpu_sab_ein_grid
  public :: ndppgpu_elastic_thinned, ndppgpu_inelastic_thinned
  public :: ndppgpu_apply_tol, ndppgpu_thin_grid
  public :: ndppgpu_sab_create, ndppgpu_sab, ndppgpu_sab_free
  public :: ndppgpu_chi
  public :: gpu_group, gpu_group_start, gpu_group_stop
  public :: ndppgpu_group_init, ndppgpu_group_unique_id, ndppgpu_group_init_rank, ndppgpu_group_ctx, ndppgpu_group_finalize
  public :: ndppgpu_group_nuclide_create, ndppgpu_group_nuclide_add_reaction, ndppgpu_group_convert_distro
  public :: ndppgpu_group_elastic, ndppgpu_group_inelastic, ndppgpu_group_nuclide_free

  ! include/ndppgpu.h: ndppgpu_params  (src/global.F90:28-59)
  type, bind(C) :: ndppgpu_params
    integer(c_int) :: scatt_type
    integer(c_int) :: order
    integer(c_int) :: mu_bins
    integer(c_int) :: nuscatter
    integer(c_int) :: ne_per_grp
    integer(c_int) :: adaptive_mu_its
    integer(c_int) :: adaptive_eout_its
    integer(c_int) :: reserved
    real(c_double) :: sab_threshold
    real(c_double) :: brent_mu_thresh
    real(c_double) :: adaptive_mu_tol
    real(c_double) :: adaptive_eout_tol
  end type ndppgpu_params

  ! include/ndppgpu.h: ndppgpu_chi_slot (one ChiData object, src/chidata_header.F90)
  type, bind(C) :: ndppgpu_chi_slot
    integer(c_int) :: law
    integer(c_int) :: delayed
    integer(c_int) :: precursor
    integer(c_int) :: threshold
    integer(c_int) :: use_pvalid
    integer(c_int) :: n_sigma
    integer(c_int) :: sigma_off
    integer(c_int) :: data_off
    integer(c_int) :: pvalid_off
    integer(c_int) :: reserved
  end type ndppgpu_chi_slot

  ! The device context of this process (one per MPI rank), created by gpu_start
  type(c_ptr), save :: gpu_ctx = c_null_ptr

  interface

    function ndppgpu_init(device, ctx) bind(C, name="ndppgpu_init") result(rc)
      import :: c_int, c_ptr
      integer(c_int), value :: device
      type(c_ptr)           :: ctx
      integer(c_int)        :: rc
    end function ndppgpu_init

    function ndppgpu_finalize(ctx) bind(C, name="ndppgpu_finalize") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int)     :: rc
    end function ndppgpu_finalize

    function ndppgpu_last_error(ctx, buf, len) bind(C, name="ndppgpu_last_error") result(n)
      import :: c_int, c_ptr, c_char
      type(c_ptr), value     :: ctx
      character(kind=c_char) :: buf(*)
      integer(c_int), value  :: len
      integer(c_int)         :: n
    end function ndppgpu_last_error

    ! type(Nuclide) scalars + energy grid + elastic xs, group structure, parameters
    function ndppgpu_nuclide_create(ctx, awr, kT, freegas_cutoff, n_grid, energy, &
         elastic_xs, e_bins, n_bins, params, nuc) &
         bind(C, name="ndppgpu_nuclide_create") result(rc)
      import :: c_int, c_ptr, c_double, ndppgpu_params
      type(c_ptr), value         :: ctx
      real(c_double), value      :: awr, kT, freegas_cutoff
      integer(c_int), value      :: n_grid, n_bins
      real(c_double), intent(in) :: energy(*), elastic_xs(*), e_bins(*)
      type(ndppgpu_params), intent(in) :: params
      type(c_ptr)                :: nuc
      integer(c_int)             :: rc
    end function ndppgpu_nuclide_create

    ! one call per ScattData slot, in the order calc_scatt fills rxn_data(:)
    ! (src/scatt.F90:88-105); array arguments are c_loc of contiguous arrays or
    ! c_null_ptr with a zero count
    function ndppgpu_nuclide_add_reaction(nuc, rxn_index, MT, Q_value, threshold, &
         scatter_in_cm, has_angle_dist, has_energy_dist, law, multiplicity, &
         yield_tab1, n_yield, sigma, n_sigma, p_valid_tab1, n_pvalid, &
         adist_energy, adist_type, adist_loc, n_adist_e, adist_data, n_adist_data, &
         edist_data, n_edist_data) &
         bind(C, name="ndppgpu_nuclide_add_reaction") result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value    :: nuc
      integer(c_int), value :: rxn_index, MT, threshold, scatter_in_cm
      integer(c_int), value :: has_angle_dist, has_energy_dist, law, multiplicity
      real(c_double), value :: Q_value
      type(c_ptr), value    :: yield_tab1, sigma, p_valid_tab1
      type(c_ptr), value    :: adist_energy, adist_type, adist_loc, adist_data, edist_data
      integer(c_int), value :: n_yield, n_sigma, n_pvalid, n_adist_e, n_adist_data, n_edist_data
      integer(c_int)        :: rc
    end function ndppgpu_nuclide_add_reaction

    function ndppgpu_convert_distro(nuc) bind(C, name="ndppgpu_convert_distro") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: nuc
      integer(c_int)     :: rc
    end function ndppgpu_convert_distro

    ! calc_elastic_grid (src/scatt.F90:603): el_mat(order, groups, NE)
    function ndppgpu_elastic(nuc, Ein, NE, el_mat) bind(C, name="ndppgpu_elastic") result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value          :: nuc
      integer(c_int), value       :: NE
      real(c_double), intent(in)  :: Ein(*)
      real(c_double), intent(out) :: el_mat(*)
      integer(c_int)              :: rc
    end function ndppgpu_elastic

    ! calc_inelastic_grid (src/scatt.F90:682); nuinel_mat = c_null_ptr unless nuscatt
    function ndppgpu_inelastic(nuc, Ein, NE, inel_mat, nuinel_mat) &
         bind(C, name="ndppgpu_inelastic") result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value          :: nuc, nuinel_mat
      integer(c_int), value       :: NE
      real(c_double), intent(in)  :: Ein(*)
      real(c_double), intent(out) :: inel_mat(*)
      integer(c_int)              :: rc
    end function ndppgpu_inelastic

    ! calc_*_grid + apply_tol_scatt + thin_grid (src/ndpp.F90:607-648); Ein is in/out
    function ndppgpu_elastic_thinned(nuc, Ein, NE, print_tol, thin_tol, tokeep, n_tokeep, &
         el_mat, n_kept, compression, max_abs_err) &
         bind(C, name="ndppgpu_elastic_thinned") result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value            :: nuc
      integer(c_int), value         :: NE, n_tokeep
      real(c_double), value         :: print_tol, thin_tol
      real(c_double), intent(inout) :: Ein(*)
      real(c_double), intent(in)    :: tokeep(*)
      real(c_double), intent(out)   :: el_mat(*)
      integer(c_int), intent(out)   :: n_kept
      real(c_double), intent(out)   :: compression, max_abs_err
      integer(c_int)                :: rc
    end function ndppgpu_elastic_thinned

    function ndppgpu_inelastic_thinned(nuc, Ein, NE, print_tol, thin_tol, tokeep, n_tokeep, &
         inel_mat, nuinel_mat, n_kept, compression, max_abs_err) &
         bind(C, name="ndppgpu_inelastic_thinned") result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value            :: nuc, nuinel_mat
      integer(c_int), value         :: NE, n_tokeep
      real(c_double), value         :: print_tol, thin_tol
      real(c_double), intent(inout) :: Ein(*)
      real(c_double), intent(in)    :: tokeep(*)
      real(c_double), intent(out)   :: inel_mat(*)
      integer(c_int), intent(out)   :: n_kept
      real(c_double), intent(out)   :: compression, max_abs_err
      integer(c_int)                :: rc
    end function ndppgpu_inelastic_thinned

    ! ndppgpu_elastic + ndppgpu_inelastic in one call (src/scatt.F90:143-150): el_mat leaves for the host while the
    ! inelastic kernels run; nuinel_mat = c_null_ptr when nuscatt is .false.
    function ndppgpu_calc_scatt(nuc, Ein_el, NE_el, el_mat, Ein_inel, NE_inel, inel_mat, nuinel_mat) &
         bind(C, name="ndppgpu_calc_scatt") result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value          :: nuc, nuinel_mat
      integer(c_int), value       :: NE_el, NE_inel
      real(c_double), intent(in)  :: Ein_el(*), Ein_inel(*)
      real(c_double), intent(out) :: el_mat(*), inel_mat(*)
      integer(c_int)              :: rc
    end function ndppgpu_calc_scatt

    ! create_Ein_grid (src/scatt.F90:166-236) on the device; lengths back, then ndppgpu_nuclide_ein_grid copies a grid
    function ndppgpu_nuclide_create_ein_grid(nuc, extend_pts, inel_extend_pts, n_el, n_inel, status) &
         bind(C, name="ndppgpu_nuclide_create_ein_grid") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value          :: nuc
      integer(c_int), value       :: extend_pts, inel_extend_pts
      integer(c_int), intent(out) :: n_el, n_inel, status
      integer(c_int)              :: rc
    end function ndppgpu_nuclide_create_ein_grid

    ! which = 0: Ein_el, 1: Ein_inel; d_Ein: c_null_ptr, or the address of a type(c_ptr) that receives the device pointer
    function ndppgpu_nuclide_ein_grid(nuc, which, Ein, d_Ein) bind(C, name="ndppgpu_nuclide_ein_grid") result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value          :: nuc, d_Ein
      integer(c_int), value       :: which
      real(c_double), intent(out) :: Ein(*)
      integer(c_int)              :: rc
    end function ndppgpu_nuclide_ein_grid

    function ndppgpu_nuclide_free(nuc) bind(C, name="ndppgpu_nuclide_free") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: nuc
      integer(c_int)     :: rc
    end function ndppgpu_nuclide_free

    ! apply_tol_scatt(data, tol) (src/scatt.F90:786) on data(L, G, NE), in place
    function ndppgpu_apply_tol(ctx, mat, NE, G, L, tol) bind(C, name="ndppgpu_apply_tol") result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value            :: ctx
      real(c_double), intent(inout) :: mat(*)
      integer(c_int), value         :: NE, G, L
      real(c_double), value         :: tol
      integer(c_int)                :: rc
    end function ndppgpu_apply_tol

    ! thin_grid (src/thin.F90:19); y2 = c_null_ptr when there is no second array
    function ndppgpu_thin_grid(ctx, x, y1, y2, NE, GL, tokeep, n_tokeep, tol, n_kept, &
         compression, max_abs_err) bind(C, name="ndppgpu_thin_grid") result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value            :: ctx, y2
      real(c_double), intent(inout) :: x(*), y1(*)
      integer(c_int), value         :: NE, GL, n_tokeep
      real(c_double), intent(in)    :: tokeep(*)
      real(c_double), value         :: tol
      integer(c_int), intent(out)   :: n_kept
      real(c_double), intent(out)   :: compression, max_abs_err
      integer(c_int)                :: rc
    end function ndppgpu_thin_grid

    ! type(SAlphaBeta) (src/ace_header.F90:201-235); Fortran column-major arrays pass
    ! unchanged; pointer arguments are c_loc(...) or c_null_ptr
    function ndppgpu_sab_create(ctx, awr, kT, threshold_inelastic, threshold_elastic, &
         n_inelastic_e_in, n_inelastic_e_out, n_inelastic_mu, secondary_mode, &
         inelastic_e_in, inelastic_sigma, inelastic_e_out, inelastic_mu, &
         cont_n_e_out, cont_e_out, cont_pdf, cont_mu, elastic_mode, n_elastic_e_in, &
         n_elastic_mu, elastic_e_in, elastic_P, elastic_mu, sab) &
         bind(C, name="ndppgpu_sab_create") result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value    :: ctx
      real(c_double), value :: awr, kT, threshold_inelastic, threshold_elastic
      integer(c_int), value :: n_inelastic_e_in, n_inelastic_e_out, n_inelastic_mu, secondary_mode
      type(c_ptr), value    :: inelastic_e_in, inelastic_sigma, inelastic_e_out, inelastic_mu
      type(c_ptr), value    :: cont_n_e_out, cont_e_out, cont_pdf, cont_mu
      integer(c_int), value :: elastic_mode, n_elastic_e_in, n_elastic_mu
      type(c_ptr), value    :: elastic_e_in, elastic_P, elastic_mu
      type(c_ptr)           :: sab
      integer(c_int)        :: rc
    end function ndppgpu_sab_create

    ! integrate_sab_el + integrate_sab_inel + combine_sab_grid (src/scatt.F90:573-591)
    function ndppgpu_sab(sab, e_bins, n_bins, scatt_type, order, Ein, NE, scatt_mat, &
         el_out, inel_out) bind(C, name="ndppgpu_sab") result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value          :: sab, el_out, inel_out
      real(c_double), intent(in)  :: e_bins(*), Ein(*)
      integer(c_int), value       :: n_bins, scatt_type, order, NE
      real(c_double), intent(out) :: scatt_mat(*)
      integer(c_int)              :: rc
    end function ndppgpu_sab

    ! sab_egrid (src/sab.F90:460-568) on the device
    function ndppgpu_sab_egrid(sab, e_bins, n_bins, sab_epts_per_bin, extend_pts, n, status) &
         bind(C, name="ndppgpu_sab_egrid") result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value          :: sab
      real(c_double), intent(in)  :: e_bins(*)
      integer(c_int), value       :: n_bins, sab_epts_per_bin, extend_pts
      integer(c_int), intent(out) :: n, status
      integer(c_int)              :: rc
    end function ndppgpu_sab_egrid

    function ndppgpu_sab_ein_grid(sab, Ein, d_Ein) bind(C, name="ndppgpu_sab_ein_grid") result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value          :: sab, d_Ein
      real(c_double), intent(out) :: Ein(*)
      integer(c_int)              :: rc
    end function ndppgpu_sab_ein_grid

    function ndppgpu_sab_free(sab) bind(C, name="ndppgpu_sab_free") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: sab
      integer(c_int)     :: rc
    end function ndppgpu_sab_free

    ! the E_in loop of calc_chi (src/chi.F90:120-153)
    function ndppgpu_chi(ctx, n_grid, energy, fission, nu_t_type, nu_t_data, n_nu_t, &
         nu_d_type, nu_d_data, n_nu_d, n_precursor, precursor_data, n_precursor_data, &
         n_slots, slots, pool, n_pool, e_bins, n_bins, Ein, NE, chi_total, chi_prompt, &
         chi_delay) bind(C, name="ndppgpu_chi") result(rc)
      import :: c_int, c_ptr, c_double, ndppgpu_chi_slot
      type(c_ptr), value          :: ctx, nu_t_data, nu_d_data, precursor_data
      integer(c_int), value       :: n_grid, nu_t_type, n_nu_t, nu_d_type, n_nu_d, n_precursor
      integer(c_int), value       :: n_precursor_data, n_slots, n_pool, n_bins, NE
      real(c_double), intent(in)  :: energy(*), fission(*), pool(*), e_bins(*), Ein(*)
      type(ndppgpu_chi_slot), intent(in) :: slots(*)
      real(c_double), intent(out) :: chi_total(*), chi_prompt(*), chi_delay(*)
      integer(c_int)              :: rc
    end function ndppgpu_chi

    ! ---- several GPUs (include/ndppgpu.h, "several GPUs"): the GPUs of this process work on one nuclide ----
    ! n_devices <= 0: every GPU of the box; devices = c_null_ptr: 0 .. n-1.  One host thread per device and
    ! ncclCommInitAll happen inside the library.
    function ndppgpu_group_init(n_devices, devices, group) bind(C, name="ndppgpu_group_init") result(rc)
      import :: c_int, c_ptr
      integer(c_int), value :: n_devices
      type(c_ptr), value    :: devices
      type(c_ptr)           :: group
      integer(c_int)        :: rc
    end function ndppgpu_group_init

    ! one GPU per MPI rank: rank 0 makes the 128-byte NCCL id, MPI_Bcast carries it, every rank joins
    function ndppgpu_group_unique_id(id128) bind(C, name="ndppgpu_group_unique_id") result(rc)
      import :: c_int, c_char
      character(kind=c_char) :: id128(128)
      integer(c_int)         :: rc
    end function ndppgpu_group_unique_id

    function ndppgpu_group_init_rank(device, rank, world, id128, group) &
         bind(C, name="ndppgpu_group_init_rank") result(rc)
      import :: c_int, c_ptr, c_char
      integer(c_int), value  :: device, rank, world
      character(kind=c_char) :: id128(128)
      type(c_ptr)            :: group
      integer(c_int)         :: rc
    end function ndppgpu_group_init_rank

    ! borrowed context of a local device (errors, statistics); 0 = the device that receives the matrices on rank 0
    function ndppgpu_group_ctx(group, local_index) bind(C, name="ndppgpu_group_ctx") result(ctx)
      import :: c_int, c_ptr
      type(c_ptr), value    :: group
      integer(c_int), value :: local_index
      type(c_ptr)           :: ctx
    end function ndppgpu_group_ctx

    function ndppgpu_group_finalize(group) bind(C, name="ndppgpu_group_finalize") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: group
      integer(c_int)     :: rc
    end function ndppgpu_group_finalize

    ! the nuclide replicated on every device of the group: arguments of ndppgpu_nuclide_create
    function ndppgpu_group_nuclide_create(group, awr, kT, freegas_cutoff, n_grid, energy, &
         elastic_xs, e_bins, n_bins, params, gnuc) &
         bind(C, name="ndppgpu_group_nuclide_create") result(rc)
      import :: c_int, c_ptr, c_double, ndppgpu_params
      type(c_ptr), value         :: group
      real(c_double), value      :: awr, kT, freegas_cutoff
      integer(c_int), value      :: n_grid, n_bins
      real(c_double), intent(in) :: energy(*), elastic_xs(*), e_bins(*)
      type(ndppgpu_params), intent(in) :: params
      type(c_ptr)                :: gnuc
      integer(c_int)             :: rc
    end function ndppgpu_group_nuclide_create

    ! arguments of ndppgpu_nuclide_add_reaction
    function ndppgpu_group_nuclide_add_reaction(gnuc, rxn_index, MT, Q_value, threshold, &
         scatter_in_cm, has_angle_dist, has_energy_dist, law, multiplicity, &
         yield_tab1, n_yield, sigma, n_sigma, p_valid_tab1, n_pvalid, &
         adist_energy, adist_type, adist_loc, n_adist_e, adist_data, n_adist_data, &
         edist_data, n_edist_data) &
         bind(C, name="ndppgpu_group_nuclide_add_reaction") result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value    :: gnuc
      integer(c_int), value :: rxn_index, MT, threshold, scatter_in_cm
      integer(c_int), value :: has_angle_dist, has_energy_dist, law, multiplicity
      real(c_double), value :: Q_value
      type(c_ptr), value    :: yield_tab1, sigma, p_valid_tab1
      type(c_ptr), value    :: adist_energy, adist_type, adist_loc, adist_data, edist_data
      integer(c_int), value :: n_yield, n_sigma, n_pvalid, n_adist_e, n_adist_data, n_edist_data
      integer(c_int)        :: rc
    end function ndppgpu_group_nuclide_add_reaction

    function ndppgpu_group_convert_distro(gnuc) bind(C, name="ndppgpu_group_convert_distro") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: gnuc
      integer(c_int)     :: rc
    end function ndppgpu_group_convert_distro

    ! calc_elastic_grid / calc_inelastic_grid sharded over the group; the matrices arrive on rank 0
    function ndppgpu_group_elastic(gnuc, Ein, NE, el_mat) bind(C, name="ndppgpu_group_elastic") result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value         :: gnuc
      real(c_double), intent(in) :: Ein(*)
      integer(c_int), value      :: NE
      real(c_double)             :: el_mat(*)
      integer(c_int)             :: rc
    end function ndppgpu_group_elastic

    function ndppgpu_group_inelastic(gnuc, Ein, NE, inel_mat, nuinel_mat) &
         bind(C, name="ndppgpu_group_inelastic") result(rc)
      import :: c_int, c_ptr, c_double
      type(c_ptr), value         :: gnuc
      real(c_double), intent(in) :: Ein(*)
      integer(c_int), value      :: NE
      real(c_double)             :: inel_mat(*)
      type(c_ptr), value         :: nuinel_mat     ! c_loc(nuinel_mat) or c_null_ptr
      integer(c_int)             :: rc
    end function ndppgpu_group_inelastic

    function ndppgpu_group_nuclide_free(gnuc) bind(C, name="ndppgpu_group_nuclide_free") result(rc)
      import :: c_int, c_ptr
      type(c_ptr), value :: gnuc
      integer(c_int)     :: rc
    end function ndppgpu_group_nuclide_free

  end interface

  type(c_ptr), save :: gpu_group = c_null_ptr     ! the device group of this process (gpu_group_start)

contains

!===============================================================================
! GPU_GROUP_START / GPU_GROUP_STOP: every GPU of the box (n_devices <= 0) or the
! first n_devices of them work together on each nuclide.  gpu_ctx becomes the
! context of the device that receives the assembled matrices, so that gpu_check,
! ndppgpu_apply_tol, ndppgpu_thin_grid and ndppgpu_sab keep working unchanged.
!===============================================================================

  subroutine gpu_group_start(n_devices)
    integer, intent(in) :: n_devices
    integer(c_int) :: rc
    rc = ndppgpu_group_init(int(n_devices, c_int), c_null_ptr, gpu_group)
    if (rc == 0) gpu_ctx = ndppgpu_group_ctx(gpu_group, 0_c_int)
    call gpu_check(rc)
  end subroutine gpu_group_start

  subroutine gpu_group_stop()
    integer(c_int) :: rc
    if (c_associated(gpu_group)) rc = ndppgpu_group_finalize(gpu_group)
    gpu_group = c_null_ptr
    gpu_ctx = c_null_ptr        ! it belonged to the group
  end subroutine gpu_group_stop

!===============================================================================
! GPU_START / GPU_STOP create and destroy the device context of this process.
! Call gpu_start from init_run (src/initialize.F90:23, after MPI_INIT, with
! device = mod(rank, number of GPUs of the node)) and gpu_stop next to
! MPI_FINALIZE (src/main.F90:29).
!===============================================================================

  subroutine gpu_start(device)
    integer, intent(in) :: device
    call gpu_check(ndppgpu_init(int(device, c_int), gpu_ctx))
  end subroutine gpu_start

  subroutine gpu_stop()
    integer(c_int) :: rc
    if (c_associated(gpu_ctx)) rc = ndppgpu_finalize(gpu_ctx)
    gpu_ctx = c_null_ptr
  end subroutine gpu_stop

!===============================================================================
! GPU_CHECK turns a non-zero status of the library into the reference's
! fatal_error (src/error.F90:79) with the library's message.
!===============================================================================

  subroutine gpu_check(rc)
    use error, only: fatal_error
    integer(c_int), intent(in) :: rc
    character(kind=c_char) :: buf(1024)
    character(1024)        :: msg
    integer                :: i, n

    if (rc == 0) return
    n = ndppgpu_last_error(gpu_ctx, buf, 1024_c_int)
    msg = ''
    do i = 1, min(n, 1023)
      msg(i:i) = buf(i)
    end do
    call fatal_error(trim(msg))
  end subroutine gpu_check

!===============================================================================
! TAB1_FLAT flattens a Tab1 into [NR, NBT(NR), INT(NR), NP, x(NP), y(NP)], the
! layout interpolate_tab1 reads (src/interpolation.F90:24-60) and the library
! expects for yield_tab1 / p_valid_tab1.
!===============================================================================

  function tab1_flat(t) result(a)
    use endf_header, only: Tab1
    type(Tab1), intent(in)      :: t
    real(c_double), allocatable :: a(:)
    integer :: nr, np

    nr = t % n_regions
    np = t % n_pairs
    allocate(a(2 + 2 * nr + 2 * np))
    a(1) = real(nr, c_double)
    if (nr > 0) then
      a(2:1 + nr)          = real(t % nbt(1:nr), c_double)
      a(2 + nr:1 + 2 * nr) = real(t % int(1:nr), c_double)
    end if
    a(2 + 2 * nr) = real(np, c_double)
    a(3 + 2 * nr:2 + 2 * nr + np)          = t % x(1:np)
    a(3 + 2 * nr + np:2 + 2 * nr + 2 * np) = t % y(1:np)
  end function tab1_flat

end module ndpp_gpu
