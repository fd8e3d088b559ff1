!> Drop-in replacement for src/dynamics/fv3jedi_lm_dynamics_mod.F90 of fv3-jedi-linearmodel.
!>
!> Same module name, same public type `fv3jedi_lm_dynamics_type`, same type-bound procedures
!> (create :69, init_nl/tl/ad :230-262, step_nl :268, step_tl :347, step_ad :460, delete :693) and the
!> same argument lists, so that src/fv3jedi_lm_mod.F90 compiles against it unchanged.  The FV3 time
!> stepping (fv_dynamics / fv_dynamics_tlm / fv_dynamics_fwd+bwd and everything below them) is
!> replaced by calls into libfv3lm_b200.so through fv3lm_b200_capi_mod; the grid, the namelist
!> handling and the domain decomposition still come from the reference's own fv_init / fv_init_pert
!> (one-time, host side), whose gridstruct arrays are uploaded to the device once.
!>
!> NOTE: this file is verified where a Fortran compiler + FMS exist (not in the build container,
!> see INTEGRATION.md); the identical C entry points are exercised from tests/ through ctypes.
module fv3jedi_lm_dynamics_mod

use iso_c_binding
use fv3jedi_lm_utils_mod
use fv3jedi_lm_kinds_mod
use fv3jedi_lm_const_mod
use fv3lm_b200_capi_mod

use mpp_mod,                only: mpp_pe, mpp_root_pe, mpp_npes, mpp_broadcast, mpp_error, FATAL
use fv_control_nlm_mod,     only: fv_init, pelist_all
use fv_control_tlmadm_mod,  only: fv_init_pert
use fv_arrays_nlm_mod,      only: fv_atmos_type, deallocate_fv_atmos_type
use fv_arrays_tlmadm_mod,   only: fv_atmos_pert_type, deallocate_fv_atmos_pert_type

implicit none
private
public :: fv3jedi_lm_dynamics_type

integer, parameter :: fvprec = 8

type fv3jedi_lm_dynamics_type
 type(fv_atmos_type),      allocatable :: FV_Atm(:)   !< grid, flags, decomposition (host side only)
 type(fv_atmos_pert_type), allocatable :: FV_AtmP(:)  !< perturbation flags (host side only)
 type(c_ptr) :: handle = c_null_ptr                    !< opaque fv3lm_handle*
 integer :: isc,iec,jsc,jec, isd,ied,jsd,jed, npz
 logical :: phis_set = .false.
 contains
  procedure :: create
  procedure :: init_nl
  procedure :: init_tl
  procedure :: init_ad
  procedure :: step_nl
  procedure :: step_tl
  procedure :: step_ad
  procedure :: delete
end type fv3jedi_lm_dynamics_type

contains

! ------------------------------------------------------------------------------

subroutine check(self, rc, what)
 class(fv3jedi_lm_dynamics_type), intent(in) :: self
 integer(c_int), intent(in) :: rc
 character(len=*), intent(in) :: what
 character(kind=c_char), pointer :: msg(:)
 character(len=512) :: text
 integer :: n
 if (rc == 0) return
 text = ''
 call c_f_pointer(fv3lm_last_error(self%handle), msg, [512])
 do n = 1, 512
    if (msg(n) == c_null_char) exit
    text(n:n) = msg(n)
 enddo
 ! the reference aborts (call exit(1), src/fv3jedi_lm_mod.F90:91-94 / mpp_error FATAL): same here
 call mpp_error(FATAL, 'fv3lm_b200 '//trim(what)//': '//trim(text))
end subroutine check

! ------------------------------------------------------------------------------

subroutine create(self,conf)

 class(fv3jedi_lm_dynamics_type), target, intent(inout) :: self
 type(fv3jedi_lm_conf), intent(inout)    :: conf

 logical, allocatable :: grids_on_this_pe(:)
 integer :: p_split = 1
 integer :: i, j
 type(fv3lm_config) :: cfg
 type(fv_atmos_type), pointer :: A
 character(kind=c_char) :: id128(128)
 integer :: id_int(128)

 ! grid, flags and decomposition exactly as the reference builds them (fv_control_nlm.F90:260, fv_control_tlmadm.F90:87)
 call fv_init(self%FV_Atm, real(conf%dt, fvprec), grids_on_this_pe, p_split)
 if (allocated(grids_on_this_pe)) deallocate(grids_on_this_pe)
 if (allocated(pelist_all)) deallocate(pelist_all)
 A => self%FV_Atm(1)
 A%ak = conf%ak; A%bk = conf%bk; A%ptop = conf%ptop
 call fv_init_pert(self%FV_Atm, self%FV_AtmP, conf%inputpert_filename)

 ! Coriolis parameter (reference create :126-139, f_coriolis_angle = 0)
 do j = A%bd%jsd, A%bd%jed+1
    do i = A%bd%isd, A%bd%ied+1
       A%gridstruct%fC(i,j) = 2.0_kind_real*omega*sin(A%gridstruct%grid(i,j,2))
    enddo
 enddo
 do j = A%bd%jsd, A%bd%jed
    do i = A%bd%isd, A%bd%ied
       A%gridstruct%f0(i,j) = 2.0_kind_real*omega*sin(A%gridstruct%agrid(i,j,2))
    enddo
 enddo

 self%isc = A%bd%isc; self%iec = A%bd%iec; self%jsc = A%bd%jsc; self%jec = A%bd%jec
 self%isd = A%bd%isd; self%ied = A%bd%ied; self%jsd = A%bd%jsd; self%jed = A%bd%jed
 self%npz = A%npz
 conf%rpe = (mpp_pe() == mpp_root_pe())

 ! flat POD configuration (fv_flags_type / fv_flags_pert_type -> fv3lm_config)
 cfg%npx = A%npx; cfg%npy = A%npy; cfg%npz = A%npz; cfg%ng = A%ng; cfg%ntiles = 6
 cfg%hydrostatic = merge(1, 0, A%flagstruct%hydrostatic)
 cfg%n_split = A%flagstruct%n_split; cfg%k_split = A%flagstruct%k_split
 cfg%nq = 4
 ! the TL/AD implement the linear schemes only: the perturbation orders are the ones that apply
 cfg%hord_mt = self%FV_AtmP(1)%flagstruct%hord_mt_pert; cfg%hord_vt = self%FV_AtmP(1)%flagstruct%hord_vt_pert
 cfg%hord_tm = self%FV_AtmP(1)%flagstruct%hord_tm_pert; cfg%hord_dp = self%FV_AtmP(1)%flagstruct%hord_dp_pert
 cfg%hord_tr = self%FV_AtmP(1)%flagstruct%hord_tr_pert
 cfg%n_sponge = self%FV_AtmP(1)%flagstruct%n_sponge_pert
 cfg%nord = self%FV_AtmP(1)%flagstruct%nord_pert
 cfg%dt = conf%dt; cfg%ptop = conf%ptop
 cfg%dddmp = self%FV_AtmP(1)%flagstruct%dddmp_pert; cfg%d2_bg = self%FV_AtmP(1)%flagstruct%d2_bg_pert
 cfg%d4_bg = self%FV_AtmP(1)%flagstruct%d4_bg_pert; cfg%vtdm4 = self%FV_AtmP(1)%flagstruct%vtdm4_pert
 cfg%d2_bg_k1 = self%FV_AtmP(1)%flagstruct%d2_bg_k1_pert; cfg%d2_bg_k2 = self%FV_AtmP(1)%flagstruct%d2_bg_k2_pert
 cfg%d_ext = A%flagstruct%d_ext; cfg%beta = A%flagstruct%beta
 cfg%zvir = zvir; cfg%kappa = kappa; cfg%cp = cp; cfg%rdgas = rgas; cfg%grav = grav
 cfg%do_vort_damp = merge(1, 0, self%FV_AtmP(1)%flagstruct%do_vort_damp_pert)
 cfg%rank = mpp_pe() - mpp_root_pe(); cfg%nranks = mpp_npes()
 cfg%layout_x = A%layout(1); cfg%layout_y = A%layout(2)
 ! one GPU per MPI rank: rank modulo the devices visible on the node (fv3lm_create calls cudaSetDevice before any allocation)
 cfg%device = -1
 ! switches the reference hands to fv_dynamics (src/dynamics/fv3jedi_lm_dynamics_mod.F90:299) whose paths are not built: refuse
 ! them here instead of computing something else silently (the ranges of beta / d_ext / a_imp are checked by fv3lm_create itself)
 if (A%flagstruct%consv_te > 0.0_kind_real) call mpp_error(FATAL, 'fv3lm_b200: consv_te > 0 (energy fixer) is not supported')
 if (A%flagstruct%consv_am) call mpp_error(FATAL, 'fv3lm_b200: consv_am (angular-momentum fixer) is not supported')
 if (A%flagstruct%fill) call mpp_error(FATAL, 'fv3lm_b200: fill (tracer filling in the remap) is not supported')
 if (A%flagstruct%tau > 0.0_kind_real) call mpp_error(FATAL, 'fv3lm_b200: tau > 0 (Rayleigh friction) is not supported')
 if (A%flagstruct%nwat /= 3) call mpp_error(FATAL, 'fv3lm_b200: nwat must be 3')
 ! q_split = 0 in fv_core_nml: tracer sub-steps chosen from the Courant numbers at run time
 cfg%q_split_dynamic = merge(1, 0, A%flagstruct%q_split == 0); cfg%q_split_max = 4
 ! two-sided mode: the fields above carry the perturbation model's switches, cfg%traj the nonlinear model's
 ! (already forced to the perturbation's by fv_control_tlmadm.F90:219-253 where split_hord / split_damp are false)
 cfg%two_sided = 1
 cfg%split_damp = merge(1, 0, self%FV_AtmP(1)%flagstruct%split_damp)
 cfg%hord_ks_pert = merge(1, 0, self%FV_AtmP(1)%flagstruct%hord_ks_pert)
 cfg%hord_ks_traj = merge(1, 0, self%FV_AtmP(1)%flagstruct%hord_ks_traj)
 cfg%d2_bg_ks = self%FV_AtmP(1)%flagstruct%d2_bg_ks_pert
 cfg%traj%hord_mt = A%flagstruct%hord_mt; cfg%traj%hord_vt = A%flagstruct%hord_vt; cfg%traj%hord_tm = A%flagstruct%hord_tm
 cfg%traj%hord_dp = A%flagstruct%hord_dp; cfg%traj%hord_tr = A%flagstruct%hord_tr
 cfg%traj%nord = A%flagstruct%nord; cfg%traj%do_vort_damp = merge(1, 0, A%flagstruct%do_vort_damp)
 cfg%traj%n_sponge = A%flagstruct%n_sponge
 cfg%traj%kord_mt = A%flagstruct%kord_mt; cfg%traj%kord_wz = A%flagstruct%kord_wz
 cfg%traj%kord_tm = A%flagstruct%kord_tm; cfg%traj%kord_tr = A%flagstruct%kord_tr
 cfg%traj%dddmp = A%flagstruct%dddmp; cfg%traj%d2_bg = A%flagstruct%d2_bg; cfg%traj%d4_bg = A%flagstruct%d4_bg
 cfg%traj%vtdm4 = A%flagstruct%vtdm4; cfg%traj%d2_bg_k1 = A%flagstruct%d2_bg_k1; cfg%traj%d2_bg_k2 = A%flagstruct%d2_bg_k2
 cfg%a_imp = A%flagstruct%a_imp; cfg%p_fac = A%flagstruct%p_fac; cfg%d_con = A%flagstruct%d_con

 call check(self, fv3lm_create(cfg, conf%ak, conf%bk, self%handle), 'create')
 fv3lm_shared_handle = self%handle      ! the physics shims (fv3lm_b200_turbulence_mod) work on the same device state

 ! NCCL bootstrap: the root PE draws the id, FMS broadcasts it
 if (mpp_npes() > 1) then
    if (conf%rpe) call check(self, fv3lm_nccl_unique_id(id128), 'nccl_unique_id')
    id_int = ichar(id128)
    call mpp_broadcast(id_int, 128, mpp_root_pe())
    id128 = char(id_int)
    call check(self, fv3lm_comm_init_nccl(self%handle, id128), 'comm_init_nccl')
 endif

 call upload_metrics(self)

endsubroutine create

! ------------------------------------------------------------------------------

!> gridstruct (model/fv_arrays_nlm.F90:115-234) -> device.  Every 2-D metric is copied into the common
!> (isd:ied+1, jsd:jed+1) shape the library uses for all staggerings.
subroutine upload_metrics(self)
 class(fv3jedi_lm_dynamics_type), target, intent(inout) :: self
 type(fv_atmos_type), pointer :: A
 real(c_double), allocatable :: buf(:,:)
 integer :: k
 character(len=8) :: nm
 A => self%FV_Atm(1)
 allocate(buf(self%isd:self%ied+1, self%jsd:self%jed+1))
#define UP2D(name, arr) buf = 0.0_c_double; buf(lbound(arr,1):ubound(arr,1), lbound(arr,2):ubound(arr,2)) = arr; call check(self, fv3lm_set_metric(self%handle, name//c_null_char, buf, 0_c_int), name)
 UP2D('area', A%gridstruct%area_64)
 UP2D('rarea', A%gridstruct%rarea)
 UP2D('area_c', A%gridstruct%area_c_64)
 UP2D('rarea_c', A%gridstruct%rarea_c)
 UP2D('dx', A%gridstruct%dx)
 UP2D('dy', A%gridstruct%dy)
 UP2D('rdx', A%gridstruct%rdx)
 UP2D('rdy', A%gridstruct%rdy)
 UP2D('dxa', A%gridstruct%dxa)
 UP2D('dya', A%gridstruct%dya)
 UP2D('rdxa', A%gridstruct%rdxa)
 UP2D('rdya', A%gridstruct%rdya)
 UP2D('dxc', A%gridstruct%dxc)
 UP2D('dyc', A%gridstruct%dyc)
 UP2D('rdxc', A%gridstruct%rdxc)
 UP2D('rdyc', A%gridstruct%rdyc)
 UP2D('cosa', A%gridstruct%cosa)
 UP2D('sina', A%gridstruct%sina)
 UP2D('rsina', A%gridstruct%rsina)
 UP2D('cosa_u', A%gridstruct%cosa_u)
 UP2D('sina_u', A%gridstruct%sina_u)
 UP2D('rsin_u', A%gridstruct%rsin_u)
 UP2D('cosa_v', A%gridstruct%cosa_v)
 UP2D('sina_v', A%gridstruct%sina_v)
 UP2D('rsin_v', A%gridstruct%rsin_v)
 UP2D('cosa_s', A%gridstruct%cosa_s)
 UP2D('rsin2', A%gridstruct%rsin2)
 UP2D('divg_u', A%gridstruct%divg_u)
 UP2D('divg_v', A%gridstruct%divg_v)
 UP2D('del6_u', A%gridstruct%del6_u)
 UP2D('del6_v', A%gridstruct%del6_v)
 UP2D('f0', A%gridstruct%f0)
 UP2D('fC', A%gridstruct%fC)
 do k = 1, 4
    write(nm, '(a,i1)') 'sin_sg', k
    UP2D(trim(nm), A%gridstruct%sin_sg(:,:,k))
    write(nm, '(a,i1)') 'cos_sg', k
    UP2D(trim(nm), A%gridstruct%cos_sg(:,:,k))
 enddo
 UP2D('agrid_lon', A%gridstruct%agrid(:,:,1))
 UP2D('agrid_lat', A%gridstruct%agrid(:,:,2))
 UP2D('grid_lon', A%gridstruct%grid(:,:,1))
 UP2D('grid_lat', A%gridstruct%grid(:,:,2))
#undef UP2D
 deallocate(buf)
 ! whole-tile 1-D edge factors, index 1..npx stored at offset ng (array position = index + ng - 1, 0-based)
 call up1d(self, 'edge_w', A%gridstruct%edge_w); call up1d(self, 'edge_e', A%gridstruct%edge_e)
 call up1d(self, 'edge_s', A%gridstruct%edge_s); call up1d(self, 'edge_n', A%gridstruct%edge_n)
 call up1d(self, 'edge_vect_w', A%gridstruct%edge_vect_w); call up1d(self, 'edge_vect_e', A%gridstruct%edge_vect_e)
 call up1d(self, 'edge_vect_s', A%gridstruct%edge_vect_s); call up1d(self, 'edge_vect_n', A%gridstruct%edge_vect_n)
 call check(self, fv3lm_set_metric_scalar(self%handle, 'da_min'//c_null_char, real(A%gridstruct%da_min, c_double)), 'da_min')
 call check(self, fv3lm_set_metric_scalar(self%handle, 'da_min_c'//c_null_char, real(A%gridstruct%da_min_c, c_double)), 'da_min_c')
 ! cubed_to_latlon coefficients (init_cubed_to_latlon): from now on step_nl also returns the A-grid winds
 call up_c2l(self, A%gridstruct%a11(self%isc:self%iec, self%jsc:self%jec), A%gridstruct%a12(self%isc:self%iec, self%jsc:self%jec), &
                   A%gridstruct%a21(self%isc:self%iec, self%jsc:self%jec), A%gridstruct%a22(self%isc:self%iec, self%jsc:self%jec))
end subroutine upload_metrics

subroutine up_c2l(self, a11, a12, a21, a22)
 class(fv3jedi_lm_dynamics_type), intent(inout) :: self
 real(kind_real), intent(in) :: a11(:,:), a12(:,:), a21(:,:), a22(:,:)      ! contiguous copies of the compute-domain sections
 call check(self, fv3lm_set_c2l(self%handle, a11, a12, a21, a22), 'set_c2l')
end subroutine up_c2l

subroutine up1d(self, name, arr)
 class(fv3jedi_lm_dynamics_type), intent(inout) :: self
 character(len=*), intent(in) :: name
 real(kind_real), intent(in) :: arr(:)
 real(c_double), allocatable :: line(:)
 integer :: ng, npx
 ng = self%FV_Atm(1)%ng; npx = self%FV_Atm(1)%npx
 allocate(line(1-ng+1 : npx+ng))     ! N + 2 ng + 1 entries, Fortran index i at position i + ng - 1
 line = 0.0_c_double
 line(lbound(arr,1):lbound(arr,1)+size(arr)-1) = arr
 call check(self, fv3lm_set_metric(self%handle, name//c_null_char, line, 1_c_int), name)
 deallocate(line)
end subroutine up1d

! ------------------------------------------------------------------------------

subroutine init_nl(self,conf,pert,traj)
 class(fv3jedi_lm_dynamics_type), intent(inout) :: self
 type(fv3jedi_lm_conf), intent(in) :: conf
 type(fv3jedi_lm_pert), intent(inout) :: pert
 type(fv3jedi_lm_traj), intent(in) :: traj
endsubroutine init_nl

subroutine init_tl(self,conf,pert,traj)
 class(fv3jedi_lm_dynamics_type), intent(inout) :: self
 type(fv3jedi_lm_conf), intent(in) :: conf
 type(fv3jedi_lm_pert), intent(inout) :: pert
 type(fv3jedi_lm_traj), intent(in) :: traj
endsubroutine init_tl

subroutine init_ad(self,conf,pert,traj)
 class(fv3jedi_lm_dynamics_type), intent(inout) :: self
 type(fv3jedi_lm_conf), intent(in) :: conf
 type(fv3jedi_lm_pert), intent(inout) :: pert
 type(fv3jedi_lm_traj), intent(in) :: traj
endsubroutine init_ad

! ------------------------------------------------------------------------------

!> the ten (isc:iec, jsc:jec, npz) members -> struct of C pointers
function traj_fields(traj, hydrostatic) result(f)
 type(fv3jedi_lm_traj), target, intent(in) :: traj
 logical, intent(in) :: hydrostatic
 type(fv3lm_fields) :: f
 f%u = c_loc(traj%u); f%v = c_loc(traj%v); f%t = c_loc(traj%t); f%delp = c_loc(traj%delp)
 f%qv = c_loc(traj%qv); f%ql = c_loc(traj%ql); f%qi = c_loc(traj%qi); f%o3 = c_loc(traj%o3)
 f%w = c_null_ptr; f%delz = c_null_ptr
 if (.not. hydrostatic) then
    f%w = c_loc(traj%w); f%delz = c_loc(traj%delz)
 endif
end function traj_fields

function pert_fields(pert, hydrostatic) result(f)
 type(fv3jedi_lm_pert), target, intent(in) :: pert
 logical, intent(in) :: hydrostatic
 type(fv3lm_fields) :: f
 f%u = c_loc(pert%u); f%v = c_loc(pert%v); f%t = c_loc(pert%t); f%delp = c_loc(pert%delp)
 f%qv = c_loc(pert%qv); f%ql = c_loc(pert%ql); f%qi = c_loc(pert%qi); f%o3 = c_loc(pert%o3)
 f%w = c_null_ptr; f%delz = c_null_ptr
 if (.not. hydrostatic) then
    f%w = c_loc(pert%w); f%delz = c_loc(pert%delz)
 endif
end function pert_fields

subroutine send_traj(self, conf, traj)
 class(fv3jedi_lm_dynamics_type), intent(inout) :: self
 type(fv3jedi_lm_conf), intent(in) :: conf
 type(fv3jedi_lm_traj), intent(in) :: traj
 ! phis is constant over the window: upload once (reference re-reads it every step, traj_to_fv3 :797-800)
 if (.not. self%phis_set) then
    call check(self, fv3lm_set_phis(self%handle, traj%phis), 'set_phis')
    self%phis_set = .true.
 endif
 ! slot = conf%n keeps the whole window's trajectory resident on the device; a host that fills the
 ! slots once per outer loop can skip this call on later inner iterations
 call check(self, fv3lm_traj_set(self%handle, int(conf%n, c_int), traj_fields(traj, conf%hydrostatic)), 'traj_set')
end subroutine send_traj

! ------------------------------------------------------------------------------

subroutine step_nl(self,conf,traj)
 class(fv3jedi_lm_dynamics_type), target, intent(inout) :: self
 type(fv3jedi_lm_traj), intent(inout) :: traj
 type(fv3jedi_lm_conf), intent(in)    :: conf
 call send_traj(self, conf, traj)
 call check(self, fv3lm_step_nl(self%handle, int(conf%n, c_int), int(conf%n+1, c_int)), 'step_nl')
 call check(self, fv3lm_traj_get(self%handle, int(conf%n+1, c_int), traj_fields(traj, conf%hydrostatic)), 'traj_get')
 call check(self, fv3lm_traj_get_winds(self%handle, traj%ua, traj%va), 'traj_get_winds')      ! reference fv3_to_traj :839-840
endsubroutine step_nl

subroutine step_tl(self,conf,traj,pert)
 class(fv3jedi_lm_dynamics_type), target, intent(inout) :: self
 type(fv3jedi_lm_conf), intent(in)    :: conf
 type(fv3jedi_lm_traj), intent(in)    :: traj
 type(fv3jedi_lm_pert), intent(inout) :: pert
 call send_traj(self, conf, traj)
 call check(self, fv3lm_step_tl(self%handle, int(conf%n, c_int), pert_fields(pert, conf%hydrostatic)), 'step_tl')
 pert%ua = 0.0_kind_real; pert%va = 0.0_kind_real       ! reference fv3_to_pert :920-921
endsubroutine step_tl

subroutine step_ad(self,conf,traj,pert)
 class(fv3jedi_lm_dynamics_type), target, intent(inout) :: self
 type(fv3jedi_lm_conf), intent(in)    :: conf
 type(fv3jedi_lm_traj), intent(in)    :: traj
 type(fv3jedi_lm_pert), intent(inout) :: pert
 call send_traj(self, conf, traj)
 call check(self, fv3lm_step_ad(self%handle, int(conf%n, c_int), pert_fields(pert, conf%hydrostatic)), 'step_ad')
 pert%ua = 0.0_kind_real; pert%va = 0.0_kind_real
endsubroutine step_ad

! ------------------------------------------------------------------------------

subroutine delete(self,conf)
 class(fv3jedi_lm_dynamics_type), intent(inout) :: self
 type(fv3jedi_lm_conf), intent(in) :: conf
 integer(c_int) :: rc
 rc = fv3lm_destroy(self%handle)
 self%handle = c_null_ptr
 fv3lm_shared_handle = c_null_ptr
 call deallocate_fv_atmos_type(self%FV_Atm(1))
 deallocate(self%FV_Atm)
 call deallocate_fv_atmos_pert_type(self%FV_AtmP(1))
 deallocate(self%FV_AtmP)
endsubroutine delete

end module fv3jedi_lm_dynamics_mod
