!> Device side of the linearised boundary-layer turbulence, called from the reference's
!> src/physics/turbulence/fv3jedi_lm_turbulence_mod.F90.  That module keeps its create (parameter table), its
!> set_ltraj up to and including the BL_DRIVER call (:375-507: the nonlinear scheme that turns the trajectory into the
!> three tridiagonal systems, once per time level) and its bookkeeping of ltraj%set; three edits hand the rest over:
!>
!>   set_ltraj :509-512   the three VTRILUPERT calls             -> call b200_turb_set_ltraj(conf, ltraj%pk, ltraj%akv, ..., ltraj%ckq)
!>   step_tl   :256-270   t2pt, seven VTRISOLVEPERT, pt2t        -> call b200_turb_step(conf, pert, 1)
!>   step_ad   :326-340   the adjoint of the same                -> call b200_turb_step(conf, pert, 2)
!>   step_nl   :187-201   the same solves on the trajectory      -> call b200_turb_step_nl(conf, traj)
!>
!> The decomposition and every solve run on the device; ltraj%pk is the one set_ltraj computed for BL_DRIVER (:443).  Not compiled in the build container (no Fortran compiler there); see INTEGRATION.md.
module fv3lm_b200_turbulence_mod

use iso_c_binding
use fv3lm_b200_capi_mod
use fv3jedi_lm_utils_mod, only: fv3jedi_lm_conf, fv3jedi_lm_pert, fv3jedi_lm_traj
use fv3jedi_lm_kinds_mod, only: kind_real
use mpp_mod,              only: mpp_error, FATAL

implicit none
private
public :: b200_turb_set_ltraj, b200_turb_step, b200_turb_step_nl

contains

subroutine check(rc, what)
 integer(c_int), intent(in) :: rc
 character(len=*), intent(in) :: what
 character(kind=c_char), pointer :: msg(:)
 character(len=512) :: text
 integer :: n
 if (rc == 0) return
 text = ''
 call c_f_pointer(fv3lm_last_error(fv3lm_shared_handle), msg, [512])
 do n = 1, 512
    if (msg(n) == c_null_char) exit
    text(n:n) = msg(n)
 enddo
 call mpp_error(FATAL, 'fv3lm_b200 '//trim(what)//': '//trim(text))
end subroutine check

!> diagonals exactly as BL_DRIVER returned them (not decomposed); slot = conf%n like the dynamics shim's trajectory slot
subroutine b200_turb_set_ltraj(conf, pk, akv, bkv, ckv, aks, bks, cks, akq, bkq, ckq)
 type(fv3jedi_lm_conf), intent(in) :: conf
 real(kind_real), target, contiguous, intent(in) :: pk(:,:,:)
 real(kind_real), target, contiguous, intent(in) :: akv(:,:,:), bkv(:,:,:), ckv(:,:,:)
 real(kind_real), target, contiguous, intent(in) :: aks(:,:,:), bks(:,:,:), cks(:,:,:)
 real(kind_real), target, contiguous, intent(in) :: akq(:,:,:), bkq(:,:,:), ckq(:,:,:)
 type(fv3lm_turb_coeffs) :: co
 co%akv = c_loc(akv); co%bkv = c_loc(bkv); co%ckv = c_loc(ckv)
 co%aks = c_loc(aks); co%bks = c_loc(bks); co%cks = c_loc(cks)
 co%akq = c_loc(akq); co%bkq = c_loc(bkq); co%ckq = c_loc(ckq)
 co%pk = c_loc(pk)           ! (c_null_ptr would make the library compute it from the delp of trajectory slot conf%n)
 co%decomposed = 0
 call check(fv3lm_turb_set_ltraj(fv3lm_shared_handle, int(conf%n, c_int), co), 'turb_set_ltraj')
end subroutine b200_turb_set_ltraj

!> phase 1 = step_tl, 2 = step_ad (the numbering of vtrisolvepert)
subroutine b200_turb_step(conf, pert, phase)
 type(fv3jedi_lm_conf), intent(in) :: conf
 type(fv3jedi_lm_pert), target, intent(inout) :: pert
 integer, intent(in) :: phase
 type(fv3lm_fields) :: f
 f%u = c_loc(pert%u); f%v = c_loc(pert%v); f%t = c_loc(pert%t); f%delp = c_loc(pert%delp)
 f%qv = c_loc(pert%qv); f%ql = c_loc(pert%ql); f%qi = c_loc(pert%qi); f%o3 = c_loc(pert%o3)
 f%w = c_null_ptr; f%delz = c_null_ptr
 if (.not. conf%hydrostatic) then
    f%w = c_loc(pert%w); f%delz = c_loc(pert%delz)
 endif
 if (phase == 1) then
    call check(fv3lm_turb_step_tl(fv3lm_shared_handle, int(conf%n, c_int), f), 'turb_step_tl')
 else
    call check(fv3lm_turb_step_ad(fv3lm_shared_handle, int(conf%n, c_int), f), 'turb_step_ad')
 endif
end subroutine b200_turb_step

!> step_nl: fv3jedi_lm_mod runs the physics after the dynamics (src/fv3jedi_lm_mod.F90:153-156).  The dynamics shim's step_nl
!> left the propagated state in slot conf%n + 1 (and in traj); the solves act on that slot with the local trajectory of conf%n,
!> then the host copy is refreshed.
subroutine b200_turb_step_nl(conf, traj)
 type(fv3jedi_lm_conf), intent(in) :: conf
 type(fv3jedi_lm_traj), target, intent(inout) :: traj
 type(fv3lm_fields) :: f
 call check(fv3lm_turb_step_nl(fv3lm_shared_handle, int(conf%n, c_int), int(conf%n+1, c_int)), 'turb_step_nl')
 f%u = c_loc(traj%u); f%v = c_loc(traj%v); f%t = c_loc(traj%t); f%delp = c_loc(traj%delp)
 f%qv = c_loc(traj%qv); f%ql = c_loc(traj%ql); f%qi = c_loc(traj%qi); f%o3 = c_loc(traj%o3)
 f%w = c_null_ptr; f%delz = c_null_ptr
 if (.not. conf%hydrostatic) then
    f%w = c_loc(traj%w); f%delz = c_loc(traj%delz)
 endif
 call check(fv3lm_traj_get(fv3lm_shared_handle, int(conf%n+1, c_int), f), 'traj_get')
end subroutine b200_turb_step_nl

end module fv3lm_b200_turbulence_mod
