!> ISO_C_BINDING interfaces of the C ABI in include/fv3lm_b200.h (libfv3lm_b200.so).
!> One interface per exported entry point the Fortran host needs; names, argument order and
!> types mirror the header one for one.  No computation happens on the Fortran side.
module fv3lm_b200_capi_mod

use iso_c_binding

implicit none
public

!> mirror of the `traj` member of `struct fv3lm_config`: the nonlinear model's switches in two-sided mode
type, bind(c) :: fv3lm_traj_flags
  integer(c_int) :: hord_mt, hord_vt, hord_tm, hord_dp, hord_tr, nord, do_vort_damp, n_sponge
  integer(c_int) :: kord_mt, kord_wz, kord_tm, kord_tr
  real(c_double) :: dddmp, d2_bg, d4_bg, vtdm4, d2_bg_k1, d2_bg_k2
end type fv3lm_traj_flags

!> mirror of `struct fv3lm_config` (include/fv3lm_b200.h); bind(c) keeps the C layout
type, bind(c) :: fv3lm_config
  integer(c_int) :: npx, npy, npz, ng, ntiles
  integer(c_int) :: hydrostatic, n_split, k_split, nq
  integer(c_int) :: hord_mt, hord_vt, hord_tm, hord_dp, hord_tr
  integer(c_int) :: n_sponge, nord
  real(c_double) :: dt, ptop
  real(c_double) :: dddmp, d2_bg, d4_bg, vtdm4, d2_bg_k1, d2_bg_k2, d_ext, beta
  real(c_double) :: zvir, kappa, cp, rdgas, grav
  integer(c_int) :: do_vort_damp
  integer(c_int) :: rank, nranks, layout_x, layout_y
  integer(c_int) :: device
  real(c_double) :: a_imp, p_fac, d_con
  integer(c_int) :: two_sided, split_damp, hord_ks_pert, hord_ks_traj
  integer(c_int) :: q_split_dynamic, q_split_max
  type(fv3lm_traj_flags) :: traj
  real(c_double) :: d2_bg_ks
end type fv3lm_config

!> mirror of `struct fv3lm_fields`: ten pointers to (isc:iec, jsc:jec, npz) REAL64 arrays
type, bind(c) :: fv3lm_fields
  type(c_ptr) :: u, v, t, delp, qv, ql, qi, o3, w, delz
end type fv3lm_fields

!> mirror of `struct fv3lm_turb_coeffs`: the diagonals BL_DRIVER returns (+ optional pk), (isc:iec, jsc:jec, npz) REAL64
type, bind(c) :: fv3lm_turb_coeffs
  type(c_ptr) :: akv, bkv, ckv, aks, bks, cks, akq, bkq, ckq, pk
  integer(c_int) :: decomposed
end type fv3lm_turb_coeffs

!> the handle created by fv3jedi_lm_dynamics_mod::create, shared with the physics shims (one handle per MPI rank / GPU)
type(c_ptr), save :: fv3lm_shared_handle = c_null_ptr

interface

  integer(c_int) function fv3lm_create(cfg, ak, bk, handle) bind(c, name='fv3lm_create')
    import :: c_int, c_double, c_ptr, fv3lm_config
    type(fv3lm_config), intent(in) :: cfg
    real(c_double), intent(in) :: ak(*), bk(*)
    type(c_ptr), intent(out) :: handle
  end function

  integer(c_int) function fv3lm_destroy(handle) bind(c, name='fv3lm_destroy')
    import :: c_int, c_ptr
    type(c_ptr), value :: handle
  end function

  type(c_ptr) function fv3lm_last_error(handle) bind(c, name='fv3lm_last_error')
    import :: c_ptr
    type(c_ptr), value :: handle
  end function

  integer(c_int) function fv3lm_set_metric(handle, name, host, is_1d) bind(c, name='fv3lm_set_metric')
    import :: c_int, c_double, c_ptr, c_char
    type(c_ptr), value :: handle
    character(kind=c_char), intent(in) :: name(*)
    real(c_double), intent(in) :: host(*)
    integer(c_int), value :: is_1d
  end function

  integer(c_int) function fv3lm_set_metric_scalar(handle, name, val) bind(c, name='fv3lm_set_metric_scalar')
    import :: c_int, c_double, c_ptr, c_char
    type(c_ptr), value :: handle
    character(kind=c_char), intent(in) :: name(*)
    real(c_double), value :: val
  end function

  integer(c_int) function fv3lm_nccl_unique_id(id128) bind(c, name='fv3lm_nccl_unique_id')
    import :: c_int, c_char
    character(kind=c_char), intent(out) :: id128(128)
  end function

  integer(c_int) function fv3lm_comm_init_nccl(handle, id128) bind(c, name='fv3lm_comm_init_nccl')
    import :: c_int, c_char, c_ptr
    type(c_ptr), value :: handle
    character(kind=c_char), intent(in) :: id128(128)
  end function

  integer(c_int) function fv3lm_set_phis(handle, phis) bind(c, name='fv3lm_set_phis')
    import :: c_int, c_double, c_ptr
    type(c_ptr), value :: handle
    real(c_double), intent(in) :: phis(*)
  end function

  integer(c_int) function fv3lm_traj_set(handle, slot, traj) bind(c, name='fv3lm_traj_set')
    import :: c_int, c_ptr, fv3lm_fields
    type(c_ptr), value :: handle
    integer(c_int), value :: slot
    type(fv3lm_fields), intent(in) :: traj
  end function

  integer(c_int) function fv3lm_traj_get(handle, slot, traj) bind(c, name='fv3lm_traj_get')
    import :: c_int, c_ptr, fv3lm_fields
    type(c_ptr), value :: handle
    integer(c_int), value :: slot
    type(fv3lm_fields), intent(in) :: traj
  end function

  integer(c_int) function fv3lm_step_nl(handle, slot_in, slot_out) bind(c, name='fv3lm_step_nl')
    import :: c_int, c_ptr
    type(c_ptr), value :: handle
    integer(c_int), value :: slot_in, slot_out
  end function

  integer(c_int) function fv3lm_step_tl(handle, slot, pert) bind(c, name='fv3lm_step_tl')
    import :: c_int, c_ptr, fv3lm_fields
    type(c_ptr), value :: handle
    integer(c_int), value :: slot
    type(fv3lm_fields), intent(in) :: pert
  end function

  integer(c_int) function fv3lm_step_ad(handle, slot, pert) bind(c, name='fv3lm_step_ad')
    import :: c_int, c_ptr, fv3lm_fields
    type(c_ptr), value :: handle
    integer(c_int), value :: slot
    type(fv3lm_fields), intent(in) :: pert
  end function

  integer(c_int) function fv3lm_pert_upload(handle, pert) bind(c, name='fv3lm_pert_upload')
    import :: c_int, c_ptr, fv3lm_fields
    type(c_ptr), value :: handle
    type(fv3lm_fields), intent(in) :: pert
  end function

  integer(c_int) function fv3lm_pert_download(handle, pert) bind(c, name='fv3lm_pert_download')
    import :: c_int, c_ptr, fv3lm_fields
    type(c_ptr), value :: handle
    type(fv3lm_fields), intent(in) :: pert
  end function

  integer(c_int) function fv3lm_step_tl_dev(handle, slot) bind(c, name='fv3lm_step_tl_dev')
    import :: c_int, c_ptr
    type(c_ptr), value :: handle
    integer(c_int), value :: slot
  end function

  integer(c_int) function fv3lm_step_ad_dev(handle, slot) bind(c, name='fv3lm_step_ad_dev')
    import :: c_int, c_ptr
    type(c_ptr), value :: handle
    integer(c_int), value :: slot
  end function

  integer(c_int) function fv3lm_set_c2l(handle, a11, a12, a21, a22) bind(c, name='fv3lm_set_c2l')
    import :: c_int, c_double, c_ptr
    type(c_ptr), value :: handle
    real(c_double), intent(in) :: a11(*), a12(*), a21(*), a22(*)
  end function

  integer(c_int) function fv3lm_traj_get_winds(handle, ua, va) bind(c, name='fv3lm_traj_get_winds')
    import :: c_int, c_double, c_ptr
    type(c_ptr), value :: handle
    real(c_double), intent(out) :: ua(*), va(*)
  end function

  integer(c_int) function fv3lm_turb_set_ltraj(handle, slot, coeffs) bind(c, name='fv3lm_turb_set_ltraj')
    import :: c_int, c_ptr, fv3lm_turb_coeffs
    type(c_ptr), value :: handle
    integer(c_int), value :: slot
    type(fv3lm_turb_coeffs), intent(in) :: coeffs
  end function

  integer(c_int) function fv3lm_turb_step_nl(handle, slot_ltraj, slot_state) bind(c, name='fv3lm_turb_step_nl')
    import :: c_int, c_ptr
    type(c_ptr), value :: handle
    integer(c_int), value :: slot_ltraj, slot_state
  end function

  integer(c_int) function fv3lm_turb_step_tl(handle, slot, pert) bind(c, name='fv3lm_turb_step_tl')
    import :: c_int, c_ptr, fv3lm_fields
    type(c_ptr), value :: handle
    integer(c_int), value :: slot
    type(fv3lm_fields), intent(in) :: pert
  end function

  integer(c_int) function fv3lm_turb_step_ad(handle, slot, pert) bind(c, name='fv3lm_turb_step_ad')
    import :: c_int, c_ptr, fv3lm_fields
    type(c_ptr), value :: handle
    integer(c_int), value :: slot
    type(fv3lm_fields), intent(in) :: pert
  end function

  integer(c_int) function fv3lm_turb_step_tl_dev(handle, slot) bind(c, name='fv3lm_turb_step_tl_dev')
    import :: c_int, c_ptr
    type(c_ptr), value :: handle
    integer(c_int), value :: slot
  end function

  integer(c_int) function fv3lm_turb_step_ad_dev(handle, slot) bind(c, name='fv3lm_turb_step_ad_dev')
    import :: c_int, c_ptr
    type(c_ptr), value :: handle
    integer(c_int), value :: slot
  end function

  integer(c_int) function fv3lm_sync(handle) bind(c, name='fv3lm_sync')
    import :: c_int, c_ptr
    type(c_ptr), value :: handle
  end function

end interface

end module fv3lm_b200_capi_mod
