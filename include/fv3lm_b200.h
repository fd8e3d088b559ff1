/* fv3lm_b200 -- C ABI of the B200-native FV3 tangent-linear / adjoint dynamics.
 *
 * This is the boundary a Fortran ISO_C_BINDING shim (see INTEGRATION.md and
 * fv3-jedi-linearmodel_b200/fortran/) binds in place of the reference's
 *   src/dynamics/fv3jedi_lm_dynamics_mod.F90   create :69, step_nl :268, step_tl :347,
 *                                              step_ad :460, delete :237
 * All functions return 0 on success, non-zero on error (message via fv3lm_last_error);
 * the reference aborts instead (src/fv3jedi_lm_mod.F90:91-94), the shim turns a non-zero
 * status into the same abort.  No C++ or torch types cross this boundary; the caller
 * owns every host array, the library owns all device memory.
 *
 * Host array layout for "halo'd" fields (module-level entry points): C order
 * [6 tiles][nk][NY][NX], NX = NY = N + 2*ng + 1, Fortran tile index (i,j) at
 * [j + ng - 1][i + ng - 1]  (one array shape serves A-, C-, D-grid and corner fields,
 * exactly like the reference's isd:ied+1 / jsd:jed+1 allocations,
 * model/fv_arrays_nlm.F90:999-1139).
 */
#ifndef FV3LM_B200_H
#define FV3LM_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct fv3lm_handle fv3lm_handle;

/* flat POD mirror of the run-time switches the path needs:
 * fv_flags_type (model/fv_arrays_nlm.F90:236-506), fv_flags_pert_type
 * (model_tlmadm/fv_arrays_tlmadm.F90:37-92) and fv3jedi_lm_conf (utils/fv3jedi_lm_utils_mod.F90:14-32) */
typedef struct fv3lm_config {
  int npx, npy, npz;      /* corner counts per tile edge (N+1) and number of layers            */
  int ng;                 /* halo width, 3 (tools/fv_mp_nlm_mod.F90:63)                        */
  int ntiles;             /* 6                                                                 */
  int hydrostatic;        /* 1 = hydrostatic                                                   */
  int n_split, k_split;   /* acoustic / remap sub-cycling                                      */
  int nq;                 /* number of tracers (4: qv ql qi o3)                                */
  int hord_mt, hord_vt, hord_tm, hord_dp, hord_tr;   /* 1, 2 or 333 (the linear schemes)      */
  int n_sponge;           /* layers using first-order (hord_*_ks = 1) transport                */
  int nord;               /* divergence damping order                                          */
  double dt;              /* model time step [s]                                               */
  double ptop;
  double dddmp, d2_bg, d4_bg, vtdm4, d2_bg_k1, d2_bg_k2, d_ext, beta;
  double zvir, kappa, cp, rdgas, grav;   /* physical constants (both constant sets, SURVEY G) */
  int do_vort_damp;
  /* domain decomposition (tools/fv_mp_nlm_mod.F90:411-441): this process is `rank` of `nranks`
   * (one process per GPU); each tile is split layout_x x layout_y (0 = choose automatically) and
   * the 6*layout_x*layout_y sub-domains are dealt out in consecutive blocks.  nranks = 0 means 1. */
  int rank, nranks, layout_x, layout_y;
  /* CUDA device of this process: 0 = keep the process's current device (the caller has called cudaSetDevice, e.g. through
   * torch); -1 = rank modulo the number of visible devices (what a Fortran/MPI host with several ranks per node wants);
   * d > 0 = device d - 1.  fv3lm_create binds it before the first allocation and every entry point re-asserts it.      */
  int device;
  /* non-hydrostatic solver (model/nh_core_nlm.F90:136-152).  0 = not set -> library default.
   * a_imp > 0.999: SIM1_solver (default 1.0); 0.5 < a_imp <= 0.999: SIM_solver in Riem_Solver3
   * (Riem_Solver_c keeps SIM1, model/nh_utils_nlm.F90:366-376); a_imp <= 0.5 (RIM_2D / SIM3) is an error.
   * p_fac: lower bound of the gas pressure relative to the hydrostatic one (default 0.05).        */
  double a_imp, p_fac;
  /* d_con > 1e-5: the kinetic energy removed by the divergence / vorticity / w damping returns as heat
   * (model/sw_core_nlm.F90:1494-1525, model/dyn_core_nlm.F90:1052-1099; needs do_vort_damp with vtdm4 > 1e-5).
   * convert_ke = F, ke_bg = 0, delt_max = 1 (model/fv_arrays_nlm.F90:312, :409-412).                */
  double d_con;
  /* Two-sided mode (the TL/AD model as fv3-jedi-linearmodel runs it, model_tlmadm/fv_arrays_tlmadm.F90:37-92): two_sided = 1
   * means the fields ABOVE (hord_*, nord, dddmp, d2_bg, d4_bg, vtdm4, do_vort_damp, d2_bg_k1, d2_bg_k2, n_sponge) are the
   * PERTURBATION model's switches (hord_*_pert, nord_pert, ..., n_sponge_pert) and `traj` holds the nonlinear model's
   * (fv_flags_type, after fv_control_tlmadm.F90:219-253 has applied split_hord / split_damp).  Per-level coefficients then
   * follow model_tlmadm/dyn_core_tlm.F90:740-926 separately for each side, and every operator whose switches differ is
   * evaluated twice -- perturbation scheme for the increment, nonlinear model's scheme for the trajectory
   * (model_tlmadm/sw_core_tlm.F90:1664-1682, 1987-1997, 2341-2366, 2436-2451).  two_sided = 0: one set of switches for both
   * (TL = exact derivative of the nonlinear step).  traj.hord_*: 1, 2, 333 or the monotone PPM schemes 8..13 of the nonlinear
   * model and its smoothness-switch schemes 3..7 (model/tp_core_nlm.F90:327-578, model/sw_core_nlm.F90:2000-2306).             */
  int two_sided, split_damp;
  int hord_ks_pert, hord_ks_traj;   /* first-order transport in the top n_sponge - 1 layers (hord_*_ks_* = 1), per side        */
  /* tracer sub-cycling.  q_split_dynamic = 0 (default): one tracer step per remap cycle (the reference's q_split = 1).
   * q_split_dynamic = 1: the reference's q_split = 0 -- the number of sub-steps follows the accumulated Courant numbers of the
   * trajectory, per level (model/fv_tracer2d_nlm.F90:351-420), up to q_split_max (0 = 3); a trajectory that needs more makes
   * fv3lm_step_* return an error.                                                                                              */
  int q_split_dynamic, q_split_max;
  struct {
    int hord_mt, hord_vt, hord_tm, hord_dp, hord_tr, nord, do_vort_damp, n_sponge;
    /* vertical remap of the trajectory (split_kord): |kord| 8 .. 14 = the limited profiles of the nonlinear model
     * (model/fv_mapz_nlm.F90:1814-2109, 2197-2463), 17 or 0 = the linear scheme; the increment always uses |kord| = 17 */
    int kord_mt, kord_wz, kord_tm, kord_tr;
    double dddmp, d2_bg, d4_bg, vtdm4, d2_bg_k1, d2_bg_k2;
  } traj;
  double d2_bg_ks;                  /* d2_bg_ks_pert: divergence damping of the remaining perturbation sponge layers           */
} fv3lm_config;

int fv3lm_create(const fv3lm_config* cfg, const double* ak, const double* bk, fv3lm_handle** out);
int fv3lm_destroy(fv3lm_handle* h);
const char* fv3lm_last_error(const fv3lm_handle* h);

/* grid metrics (gridstruct members, model/fv_arrays_nlm.F90:115-234). 2-D arrays are
 * [6][NY][NX]; the 1-D edge_* factors are [6][NX]; sin_sg/cos_sg are passed per index
 * as "sin_sg1".."sin_sg4", "cos_sg1".."cos_sg4". */
int fv3lm_set_metric(fv3lm_handle* h, const char* name, const double* host, int is_1d);
int fv3lm_set_metric_scalar(fv3lm_handle* h, const char* name, double value);   /* da_min, da_min_c */

/* Kernel-family entry points used by the parity tests: run one family ("module") of the
 * hot path in mode 0 = NL, 1 = TL, 2 = AD on host arrays.
 *   traj[f] : in/out trajectory array of field names[f]
 *   pert[f] : TL: perturbation in (inputs) / out (outputs);
 *             AD: adjoint in (outputs) / out (inputs).  NULL = field not active.
 * Modules and their field names are listed by fv3lm_module_list(). */
int fv3lm_module_run(fv3lm_handle* h, const char* module, int mode, int nfields, const char* const* names,
                     double* const* traj, double* const* pert, int nparams, const char* const* pnames,
                     const double* pvals);
const char* fv3lm_module_list(void);

/* ---- step-level API: the drop-in for fv3jedi_lm_dynamics_mod step_nl/step_tl/step_ad ----------
 * (src/dynamics/fv3jedi_lm_dynamics_mod.F90:268, :347, :460).  Field arrays are the compute
 * domain only, (isc:iec, jsc:jec, npz) per tile in Fortran order = C order [tile][k][j][i],
 * exactly the members of fv3jedi_lm_traj / fv3jedi_lm_pert (utils/fv3jedi_lm_utils_mod.F90:35-54).
 * w, delz are read only when the handle is non-hydrostatic.  The trajectory is kept in a
 * device-resident window indexed by `slot` (one slot per time level of the assimilation window). */
typedef struct fv3lm_fields { double *u, *v, *t, *delp, *qv, *ql, *qi, *o3, *w, *delz; } fv3lm_fields;
int fv3lm_set_phis(fv3lm_handle* h, const double* phis);                       /* [tile][j][i] */
int fv3lm_traj_set(fv3lm_handle* h, int slot, const fv3lm_fields* traj);       /* host -> window slot */
int fv3lm_traj_get(fv3lm_handle* h, int slot, fv3lm_fields* traj);
int fv3lm_step_nl(fv3lm_handle* h, int slot_in, int slot_out);                 /* slot_out = N(slot_in), on device */
/* A-grid lon / lat winds of the propagated trajectory: cubed_to_latlon with c2l_ord = 4 at the end of fv_dynamics
 * (model/fv_dynamics_nlm.F90:738, model/fv_grid_utils_nlm.F90:2334-2472), returned by the reference's step_nl in traj%ua, traj%va
 * (src/dynamics/fv3jedi_lm_dynamics_mod.F90:839-840).  fv3lm_set_c2l takes gridstruct%a11 .. a22 (compact 2-D); from then on every
 * fv3lm_step_nl also computes them, fv3lm_traj_get_winds copies those of the last step to the host.                            */
int fv3lm_set_c2l(fv3lm_handle* h, const double* a11, const double* a12, const double* a21, const double* a22);
int fv3lm_traj_get_winds(fv3lm_handle* h, double* ua, double* va);
int fv3lm_step_tl(fv3lm_handle* h, int slot, fv3lm_fields* pert);              /* pert in/out on the host */
int fv3lm_step_ad(fv3lm_handle* h, int slot, fv3lm_fields* pert);              /* adjoint variables in/out */
/* device-resident increments: a whole window without PCIe round trips */
int fv3lm_pert_upload(fv3lm_handle* h, const fv3lm_fields* pert);
int fv3lm_pert_download(fv3lm_handle* h, fv3lm_fields* pert);
int fv3lm_step_tl_dev(fv3lm_handle* h, int slot);
int fv3lm_step_ad_dev(fv3lm_handle* h, int slot);
/* ---- linearised boundary-layer turbulence: replaces step_nl / step_tl / step_ad of
 * src/physics/turbulence/fv3jedi_lm_turbulence_mod.F90 (:149 / :218 / :285) and the device part of set_ltraj (:375-533).
 * BL_DRIVER (bldriver.F90), which turns the trajectory into the three tridiagonal systems once per time level, stays with the
 * caller; its output goes in here.  Arrays are compact (isc:iec, jsc:jec, npz) like fv3lm_fields.                             */
typedef struct fv3lm_turb_coeffs {
  const double *akv, *bkv, *ckv;   /* lower / main / upper diagonal for the winds                                   (:494-496) */
  const double *aks, *bks, *cks;   /* ... for potential temperature                                                            */
  const double *akq, *bkq, *ckq;   /* ... for specific humidity and the tracers qi, ql, o3                                     */
  const double *pk;                /* p^kappa at mid levels (ltraj%pk); NULL: computed on the device from the delp of the
                                      trajectory slot, fv3lm_config.ptop and .kappa (utils/fv3jedi_lm_utils_mod.F90:359-391)   */
  int decomposed;                  /* 0: diagonals as BL_DRIVER returns them, VTRILUPERT (:562-579) runs on the device;
                                      1: already LU-decomposed by the caller                                                   */
} fv3lm_turb_coeffs;
int fv3lm_turb_set_ltraj(fv3lm_handle* h, int slot, const fv3lm_turb_coeffs* coeffs);
int fv3lm_turb_step_nl(fv3lm_handle* h, int slot_ltraj, int slot_state);          /* trajectory fields of slot_state, in place */
int fv3lm_turb_step_tl(fv3lm_handle* h, int slot, fv3lm_fields* pert);            /* pert in/out on the host; u v t qv qi ql o3 change */
int fv3lm_turb_step_ad(fv3lm_handle* h, int slot, fv3lm_fields* pert);
int fv3lm_turb_step_tl_dev(fv3lm_handle* h, int slot);                            /* on the device-resident increments */
int fv3lm_turb_step_ad_dev(fv3lm_handle* h, int slot);
/* bench helpers: CUDA-event timing of TL+AD steps on resident data; program statistics */
int fv3lm_time_steps(fv3lm_handle* h, int slot, int warmup, int iters, double* ms_tl_ad);
int fv3lm_time_turb(fv3lm_handle* h, int slot, int warmup, int iters, double* ms_tl_ad);   /* turbulence solves, same convention */
int fv3lm_program_stats(fv3lm_handle* h, const char* module, double* out4);
int fv3lm_profile_steps(fv3lm_handle* h, int slot, int iters, char* buf, int buflen);   /* per-op event times, text */

/* ---- domain decomposition and inter-rank halo exchange (one process per GPU) -------------------
 * Replaces FMS mpp_define_mosaic / mpp_update_domains(_ad) / mpp_get_boundary(_ad)
 * (tools/fv_mp_nlm_mod.F90:285-591, model_tlmadm/fv_mp_tlm.F90:420-852, fv_mp_adm.F90:488-725).
 * With cfg.nranks > 1 every array of the API holds only this rank's sub-domains:
 *   halo'd fields  [nsub][nk][nyl + 7][nxl + 7],  step-level fields [nsub][nk][nyl][nxl],
 * sub-domain l of this rank = tile[l], tile-global origin (i0[l], j0[l]) from fv3lm_decomp_info.
 * Rank 0 obtains a 128-byte NCCL id with fv3lm_nccl_unique_id, the host broadcasts it (MPI / FMS in
 * the Fortran host), every rank passes it to fv3lm_comm_init_nccl. */
int fv3lm_decomp_info(const fv3lm_handle* h, int* out6, int* tile, int* i0, int* j0);
int fv3lm_nccl_unique_id(char* out128);
int fv3lm_comm_init_nccl(fv3lm_handle* h, const char* id128);
int fv3lm_comm_stats(const fv3lm_handle* h, double* out2);   /* exchanges issued, bytes sent */
/* TEST-ONLY transport of the host-emulation build (tests/ drive it with torch.distributed gloo) */
typedef void (*fv3lm_exchange_fn)(void* user, int npeers, const int* peers, double* const* sendbuf, const size_t* sendcount,
                                  double* const* recvbuf, const size_t* recvcount);
int fv3lm_comm_set_callback(fv3lm_handle* h, fv3lm_exchange_fn fn, void* user);

/* counters for bench.py */
long long fv3lm_launch_count(void);
double fv3lm_pool_peak_bytes(const fv3lm_handle* h);
int fv3lm_sync(fv3lm_handle* h);

#ifdef __cplusplus
}
#endif
#endif
