// dyn_core / fv_dynamics program builders
#pragma once
#include "engine.h"
#include "mosaic.h"
#include "stages_tp.h"
#include "stages_csw.h"
#include "stages_dsw.h"
#include "stages_dyn.h"

namespace fv3lm {

struct DynConfig {
  bool hydrostatic = true;
  int n_split = 1, k_split = 1, nq = 0;
  double bdt = 900.0;      // time step of one dyn_core call (dt / k_split)
  int nord = 1, hord_mt = 2, hord_vt = 2, hord_tm = 2, hord_dp = 2, hord_tr = 2;
  int n_sponge = 0, n_sponge_ord = 0;
  double d2_bg = 0.015, d2_bg_k1 = 4.0, d2_bg_k2 = 2.0, d4_bg = 0.15, dddmp = 0.2, vtdm4 = 0.0005;
  bool do_vort_damp = true;
  double ptop = 1.0, akap = 2.0 / 7.0, cp_air = 1004.6, rdgas = 287.05, grav = 9.80665, zvir = 0.6078;
  double a_imp = 1.0, p_fac = 0.05;
  double d_ext = 0.0;      // > 0 (hydrostatic only): external-mode divergence damping (model/dyn_core_nlm.F90:642-726, one_grad_p :1713-1727); the non-hydrostatic gradient never reads it
  double beta = 0.0;       // > 0: grad1_p_update / split_p_grad instead of one_grad_p / nh_p_grad (model/dyn_core_nlm.F90:865-876)
  int q_split = 1, q_split_max = 3;   // tracer sub-cycling: 1 = one step; 0 = chosen at run time from the Courant numbers (fv_tracer2d_nlm.F90:351-420)
  int kord_mt = 17, kord_wz = 17, kord_tm = 17, kord_tr = 17;   // trajectory remap (two-sided mode: 8 .. 14 monotone); the increment is always |kord| = 17
  double d_con = 0.0, delt_max = 1.0;   // dissipative heating (model/fv_arrays_nlm.F90:409-411); convert_ke = F, ke_bg = 0
  // Two-sided mode: the members above are the nonlinear model's switches (trajectory), `pert` the perturbation model's
  // (model_tlmadm/fv_arrays_tlmadm.F90:37-92)
  struct PertSide {
    bool on = false, split_damp = false, hord_ks_pert = true, hord_ks_traj = true, do_vort_damp = true;
    int hord_mt = 2, hord_vt = 2, hord_tm = 2, hord_dp = 2, hord_tr = 2, nord = 1, n_sponge = 0;
    double d2_bg = 0.015, d2_bg_k1 = 4.0, d2_bg_k2 = 2.0, d2_bg_ks = 2.0, d4_bg = 0.15, dddmp = 0.2, vtdm4 = 0.0005;
  } pert;
};

struct DynState { int u, v, w, delz, pt, delp, phis; };
struct DynOut { int u, v, w, delz, pt, delp, mfx, mfy, cx, cy, pkz, pe, peln, pk, ws; };

void level_params(const DynConfig& c, int K, DswParams& d);
// perturbation-side per-level switches (model_tlmadm/dyn_core_tlm.F90:835-921); false when the configuration is one-sided
bool level_params_pert(const DynConfig& c, int K, DswParams& d);
// del2_cubed(q, cd, nmax): returns the filtered field (q's halo is exchanged in place first)
int build_del2_cubed(Program& P, Mosaic& mo, int q, double cd, int nmax, int nk, const std::string& tag);
// end of dyn_core (model/dyn_core_nlm.F90:1052-1099): filter the accumulated heat source, add it to pt.  aux = pkz (hydrostatic) or delz
int build_heat_update(Program& P, Mosaic& mo, const DynConfig& c, int heat, int pt, int delp, int aux, const std::string& tag);
DynOut build_dyn_core(Program& P, Mosaic& mo, const DynConfig& c, DynState s, const std::string& tag = "dyn");
struct ModuleParams;
void dyn_config_from(DynConfig& c, const ModuleParams& prm);

}  // namespace fv3lm
