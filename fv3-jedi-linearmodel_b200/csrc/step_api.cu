// Step-level C ABI: the drop-in for fv3jedi_lm_dynamics_mod's step_nl / step_tl / step_ad
// (src/dynamics/fv3jedi_lm_dynamics_mod.F90:268/:347/:460) with a device-resident trajectory
// window and device-resident perturbation state.
#include "../../include/fv3lm_b200.h"
#include "capi_internal.h"
#include "fvdyn.h"
#include "modules.h"

using namespace fv3lm;

namespace {

const char* kFieldNames[NFIELD] = {"u", "v", "t", "delp", "qv", "ql", "qi", "o3", "w", "delz"};

// compact [tile][k][N][N]  <->  halo'd [tile][k][NY][pitch]
struct KExpand {
  Geom g; const double* c; double* h; int nk;
  DEV void operator()(int ii, int jj, int z) const {
    int i = ii - (g.ng - 1), j = jj - (g.ng - 1);
    double v = 0.0;
    if (i >= 1 && i <= g.ie && j >= 1 && j <= g.je) v = c[((size_t)z * g.je + (j - 1)) * g.ie + (i - 1)];
    h[(size_t)z * g.slab + (size_t)jj * g.pitch + ii] = v;
  }
};
struct KCompact {
  Geom g; double* c; const double* h; int nk;
  DEV void operator()(int ii, int jj, int z) const {
    int i = ii - (g.ng - 1), j = jj - (g.ng - 1);
    if (i >= 1 && i <= g.ie && j >= 1 && j <= g.je) c[((size_t)z * g.je + (j - 1)) * g.ie + (i - 1)] = h[(size_t)z * g.slab + (size_t)jj * g.pitch + ii];
  }
};

size_t compact_doubles(const Geom& g, int nk) { return (size_t)g.ntile * nk * g.ie * g.je; }

void ensure_runner(fv3lm_handle* h) {
  if (h->step) return;
  if (h->ak.empty()) throw std::runtime_error("fv3lm: ak/bk were not given to fv3lm_create");
  std::unique_ptr<StepRunner> owner(new StepRunner());     // released into the handle only once it is completely built
  StepRunner* r = owner.get();
  const Geom& g = h->dv.g;
  r->P.dv = &h->dv; r->P.name = "step";
  ModuleParams prm; prm.cfg = &h->cfg; prm.ak = &h->ak; prm.bk = &h->bk;
  prm.v["bdt"] = h->cfg.dt;
  build_module("step", r->P, h->mo, r->io, prm);
  r->nf = h->cfg.hydrostatic ? 8 : 10;
  for (auto& kv : r->io.inputs) {
    Value& v = r->P.vals[kv.second];
    v.traj = (double*)dev::alloc(r->P.val_doubles(kv.second) * sizeof(double));
    dev::zero(v.traj, r->P.val_doubles(kv.second) * sizeof(double));
    v.pert = (double*)dev::alloc(r->P.val_doubles(kv.second) * sizeof(double));
    dev::zero(v.pert, r->P.val_doubles(kv.second) * sizeof(double));
    r->in_id[kv.first] = kv.second;
  }
  for (auto& kv : r->io.outputs) {
    Value& v = r->P.vals[kv.second];
    if (!v.traj) { v.traj = (double*)dev::alloc(r->P.val_doubles(kv.second) * sizeof(double)); dev::zero(v.traj, r->P.val_doubles(kv.second) * sizeof(double)); }
    if (!v.pert) { v.pert = (double*)dev::alloc(r->P.val_doubles(kv.second) * sizeof(double)); dev::zero(v.pert, r->P.val_doubles(kv.second) * sizeof(double)); }
    r->out_id[kv.first] = kv.second;
  }
  for (int f = 0; f < NFIELD; f++) { r->pert[f] = nullptr; }
  for (int f = 0; f < r->nf; f++) r->pert[f] = (double*)dev::alloc(compact_doubles(g, g.K) * sizeof(double));
  r->phis = (double*)dev::alloc(compact_doubles(g, 1) * sizeof(double));
  dev::zero(r->phis, compact_doubles(g, 1) * sizeof(double));
  h->step = owner.release();
}

double** field_ptr(fv3lm_fields* f, int n) {
  switch (n) {
    case 0: return &f->u; case 1: return &f->v; case 2: return &f->t; case 3: return &f->delp; case 4: return &f->qv;
    case 5: return &f->ql; case 6: return &f->qi; case 7: return &f->o3; case 8: return &f->w; default: return &f->delz;
  }
}

double* slot_field(fv3lm_handle* h, int slot, int f) {
  StepRunner* r = h->step;
  const Geom& g = h->dv.g;
  if (slot < 0) throw std::runtime_error("fv3lm: negative trajectory slot");
  if ((int)r->slots.size() <= slot) r->slots.resize(slot + 1);
  auto& s = r->slots[slot];
  if (s.empty()) { s.assign(NFIELD, nullptr); for (int n = 0; n < r->nf; n++) s[n] = (double*)dev::alloc(compact_doubles(g, g.K) * sizeof(double)); }
  return s[f];
}

void expand(fv3lm_handle* h, const double* c, double* hal, int nk) { const Geom& g = h->dv.g; launch3d(KExpand{g, c, hal, nk}, g.NX, g.NY, g.ntile * nk); }
void compact(fv3lm_handle* h, double* c, const double* hal, int nk) { const Geom& g = h->dv.g; launch3d(KCompact{g, c, hal, nk}, g.NX, g.NY, g.ntile * nk); }

// load the trajectory of `slot` (and phis) into the program's input arrays
void load_traj(fv3lm_handle* h, int slot) {
  StepRunner* r = h->step; const Geom& g = h->dv.g;
  for (int f = 0; f < NFIELD; f++) {
    auto it = r->in_id.find(kFieldNames[f]);
    if (it == r->in_id.end()) continue;
    Value& v = r->P.vals[it->second];
    if (f < r->nf) expand(h, slot_field(h, slot, f), v.traj, g.K);
    else dev::zero(v.traj, r->P.val_doubles(it->second) * sizeof(double));
  }
  expand(h, r->phis, r->P.vals[r->in_id["phis"]].traj, 1);
}

// run the program eagerly, or (product build, default) through a captured CUDA graph
void run_program(fv3lm_handle* h, int mode) {
  StepRunner* r = h->step;
#ifndef FV3LM_HOST_EMU
  static const bool no_graph = getenv("FV3LM_NO_GRAPH") != nullptr || getenv("FV3LM_SYNC_CHECK") != nullptr;
  StepGraph& sg = r->graph[mode];
  if (no_graph || dev::profiling || mode == MODE_NL) { r->P.run((Mode)mode); return; }
  if (sg.state == 0) { r->P.run((Mode)mode); sg.state = 1; return; }          // eager: lets the pool reach its peak
  if (sg.state == 1) {
    // identical second run, recorded instead of executed: the pool hands out the same buffers in the same order
    const long long l0 = dev::launches, e0 = h->comm.n_exchanges; const double b0 = h->comm.bytes_sent;
    const size_t pool0 = h->dv.pool.bytes_total;
    cudaGraph_t g = nullptr;
    if (cudaStreamBeginCapture(dev::stream(), cudaStreamCaptureModeThreadLocal) != cudaSuccess) throw std::runtime_error("fv3lm: cudaStreamBeginCapture failed");
    try { r->P.run((Mode)mode); }
    catch (...) { cudaStreamEndCapture(dev::stream(), &g); if (g) cudaGraphDestroy(g); throw; }
    if (cudaStreamEndCapture(dev::stream(), &g) != cudaSuccess || !g) throw std::runtime_error("fv3lm: stream capture of the step program failed");
    if (h->dv.pool.bytes_total != pool0) { cudaGraphDestroy(g); throw std::runtime_error("fv3lm: device pool grew during graph capture"); }
    cudaGraphExec_t ex = nullptr;
    if (cudaGraphInstantiate(&ex, g, 0) != cudaSuccess) { cudaGraphDestroy(g); throw std::runtime_error("fv3lm: cudaGraphInstantiate failed"); }
    cudaGraphDestroy(g);
    sg.exec = ex; sg.launches = dev::launches - l0; sg.exchanges = h->comm.n_exchanges - e0; sg.bytes_sent = h->comm.bytes_sent - b0;
    sg.state = 2;
    dev::launches = l0; h->comm.n_exchanges = e0; h->comm.bytes_sent = b0;     // nothing ran yet: the replay below counts
  }
  if (cudaGraphLaunch((cudaGraphExec_t)sg.exec, dev::stream()) != cudaSuccess) throw std::runtime_error("fv3lm: cudaGraphLaunch failed");
  dev::launches += sg.launches; h->comm.n_exchanges += sg.exchanges; h->comm.bytes_sent += sg.bytes_sent;
#else
  r->P.run((Mode)mode);
#endif
}

void run_step(fv3lm_handle* h, int slot, int mode) {
  StepRunner* r = h->step; const Geom& g = h->dv.g;
  load_traj(h, slot);
  // activity pattern is fixed per mode (the captured graphs depend on it)
  for (auto& kv : r->in_id) r->P.vals[kv.second].active = false;
  if (mode == MODE_TL) {
    for (int f = 0; f < r->nf; f++) { Value& v = r->P.vals[r->in_id[kFieldNames[f]]]; v.active = true; expand(h, r->pert[f], v.pert, g.K); }
    run_program(h, MODE_TL);
    for (int f = 0; f < r->nf; f++) compact(h, r->pert[f], r->P.vals[r->out_id[std::string(kFieldNames[f]) + "_n"]].pert, g.K);
  } else if (mode == MODE_AD) {
    for (int f = 0; f < r->nf; f++) {
      Value& vi = r->P.vals[r->in_id[kFieldNames[f]]]; vi.active = true;
      dev::zero(vi.pert, r->P.val_doubles(r->in_id[kFieldNames[f]]) * sizeof(double));
      expand(h, r->pert[f], r->P.vals[r->out_id[std::string(kFieldNames[f]) + "_n"]].pert, g.K);
    }
    run_program(h, MODE_AD);
    for (int f = 0; f < r->nf; f++) compact(h, r->pert[f], r->P.vals[r->in_id[kFieldNames[f]]].pert, g.K);
  } else {
    run_program(h, MODE_NL);
  }
  if (!r->P.status_flags.empty()) r->P.check_status_flags();     // (one small device-to-host read; only programs with run-time checks)
}

}  // namespace

extern "C" {

int fv3lm_set_phis(fv3lm_handle* h, const double* phis) {
  FV3LM_TRY_H(h)
  ensure_runner(h);
  dev::h2d(h->step->phis, phis, compact_doubles(h->dv.g, 1) * sizeof(double));
  dev::sync();
  FV3LM_CATCH(h)
}

int fv3lm_traj_set(fv3lm_handle* h, int slot, const fv3lm_fields* traj) {
  FV3LM_TRY_H(h)
  ensure_runner(h);
  for (int f = 0; f < h->step->nf; f++) {
    double* src = *field_ptr(const_cast<fv3lm_fields*>(traj), f);
    if (!src) throw std::runtime_error(std::string("fv3lm_traj_set: missing field ") + kFieldNames[f]);
    dev::h2d(slot_field(h, slot, f), src, compact_doubles(h->dv.g, h->dv.g.K) * sizeof(double));
  }
  dev::sync();
  FV3LM_CATCH(h)
}

int fv3lm_traj_get(fv3lm_handle* h, int slot, fv3lm_fields* traj) {
  FV3LM_TRY_H(h)
  ensure_runner(h);
  for (int f = 0; f < h->step->nf; f++) { double* dst = *field_ptr(traj, f); if (dst) dev::d2h(dst, slot_field(h, slot, f), compact_doubles(h->dv.g, h->dv.g.K) * sizeof(double)); }
  FV3LM_CATCH(h)
}

int fv3lm_step_nl(fv3lm_handle* h, int slot_in, int slot_out) {
  FV3LM_TRY_H(h)
  ensure_runner(h);
  run_step(h, slot_in, MODE_NL);
  StepRunner* r = h->step;
  for (int f = 0; f < r->nf; f++) compact(h, slot_field(h, slot_out, f), r->P.vals[r->out_id[std::string(kFieldNames[f]) + "_n"]].traj, h->dv.g.K);
  if (r->c2l) {
    // cubed_to_latlon at the end of fv_dynamics (model/fv_dynamics_nlm.F90:738): the program reads the step's u / v output arrays in
    // place (rows je + 1 / columns ie + 1 are part of them; its halo update only touches their halo)
    C2lRunner* c = r->c2l;
    c->P.vals[c->id["u"]].traj = r->P.vals[r->out_id["u_n"]].traj;
    c->P.vals[c->id["v"]].traj = r->P.vals[r->out_id["v_n"]].traj;
    for (auto& kv : c->io.inputs) c->P.vals[kv.second].active = false;
    c->P.run(MODE_NL);
    compact(h, c->ua, c->P.vals[c->id["ua"]].traj, h->dv.g.K);
    compact(h, c->va, c->P.vals[c->id["va"]].traj, h->dv.g.K);
  }
  dev::sync();
  FV3LM_CATCH(h)
}

// gridstruct%a11 .. a22 (init_cubed_to_latlon, model/fv_grid_utils_nlm.F90:2248-2310), compact (isc:iec, jsc:jec).  Once set, every
// fv3lm_step_nl also produces the A-grid lon / lat winds of its result (fv3jedi_lm_dynamics_mod.F90:839-840: traj%ua, traj%va).
int fv3lm_set_c2l(fv3lm_handle* h, const double* a11, const double* a12, const double* a21, const double* a22) {
  FV3LM_TRY_H(h)
  ensure_runner(h);
  StepRunner* r = h->step; const Geom& g = h->dv.g;
  if (!a11 || !a12 || !a21 || !a22) throw std::runtime_error("fv3lm_set_c2l: null array");
  if (!r->c2l) {
    auto* c = new C2lRunner();
    c->P.dv = &h->dv; c->P.name = "c2l";
    ModuleParams prm; prm.cfg = &h->cfg; prm.ak = &h->ak; prm.bk = &h->bk;
    build_module("c2l_ord4", c->P, h->mo, c->io, prm);
    for (auto& kv : c->io.inputs) c->id[kv.first] = kv.second;
    for (auto& kv : c->io.outputs) c->id[kv.first] = kv.second;
    for (const char* n : {"a11", "a12", "a21", "a22", "ua", "va"}) {
      const int id = c->id[n];
      c->P.vals[id].traj = (double*)dev::alloc(c->P.val_doubles(id) * sizeof(double));
      dev::zero(c->P.vals[id].traj, c->P.val_doubles(id) * sizeof(double));
    }
    c->ua = (double*)dev::alloc(compact_doubles(g, g.K) * sizeof(double));
    c->va = (double*)dev::alloc(compact_doubles(g, g.K) * sizeof(double));
    dev::zero(c->ua, compact_doubles(g, g.K) * sizeof(double));
    dev::zero(c->va, compact_doubles(g, g.K) * sizeof(double));
    r->c2l = c;
  }
  double* tmp = (double*)dev::alloc(compact_doubles(g, 1) * sizeof(double));
  const double* src[4] = {a11, a12, a21, a22};
  static const char* nm[4] = {"a11", "a12", "a21", "a22"};
  for (int n = 0; n < 4; n++) {
    dev::h2d(tmp, src[n], compact_doubles(g, 1) * sizeof(double));
    expand(h, tmp, r->c2l->P.vals[r->c2l->id[nm[n]]].traj, 1);
  }
  dev::sync();
  dev::free_(tmp);
  FV3LM_CATCH(h)
}

// ua, va of the last fv3lm_step_nl (compact, like the fields)
int fv3lm_traj_get_winds(fv3lm_handle* h, double* ua, double* va) {
  FV3LM_TRY_H(h)
  ensure_runner(h);
  StepRunner* r = h->step; const Geom& g = h->dv.g;
  if (!r->c2l) throw std::runtime_error("fv3lm_traj_get_winds: fv3lm_set_c2l was not called");
  if (ua) dev::d2h(ua, r->c2l->ua, compact_doubles(g, g.K) * sizeof(double));
  if (va) dev::d2h(va, r->c2l->va, compact_doubles(g, g.K) * sizeof(double));
  FV3LM_CATCH(h)
}

int fv3lm_pert_upload(fv3lm_handle* h, const fv3lm_fields* pert) {
  FV3LM_TRY_H(h)
  ensure_runner(h);
  for (int f = 0; f < h->step->nf; f++) {
    double* src = *field_ptr(const_cast<fv3lm_fields*>(pert), f);
    if (!src) throw std::runtime_error(std::string("fv3lm_pert_upload: missing field ") + kFieldNames[f]);
    dev::h2d(h->step->pert[f], src, compact_doubles(h->dv.g, h->dv.g.K) * sizeof(double));
  }
  FV3LM_CATCH(h)
}

int fv3lm_pert_download(fv3lm_handle* h, fv3lm_fields* pert) {
  FV3LM_TRY_H(h)
  ensure_runner(h);
  for (int f = 0; f < h->step->nf; f++) { double* dst = *field_ptr(pert, f); if (dst) dev::d2h(dst, h->step->pert[f], compact_doubles(h->dv.g, h->dv.g.K) * sizeof(double)); }
  FV3LM_CATCH(h)
}

int fv3lm_step_tl_dev(fv3lm_handle* h, int slot) {
  FV3LM_TRY_H(h)
  ensure_runner(h);
  run_step(h, slot, MODE_TL);
  FV3LM_CATCH(h)
}

int fv3lm_step_ad_dev(fv3lm_handle* h, int slot) {
  FV3LM_TRY_H(h)
  ensure_runner(h);
  run_step(h, slot, MODE_AD);
  FV3LM_CATCH(h)
}

int fv3lm_step_tl(fv3lm_handle* h, int slot, fv3lm_fields* pert) {
  int rc = fv3lm_pert_upload(h, pert);
  if (!rc) rc = fv3lm_step_tl_dev(h, slot);
  if (!rc) rc = fv3lm_pert_download(h, pert);
  return rc;
}

int fv3lm_step_ad(fv3lm_handle* h, int slot, fv3lm_fields* pert) {
  int rc = fv3lm_pert_upload(h, pert);
  if (!rc) rc = fv3lm_step_ad_dev(h, slot);
  if (!rc) rc = fv3lm_pert_download(h, pert);
  return rc;
}

// ---- linearised boundary-layer turbulence (src/physics/turbulence/fv3jedi_lm_turbulence_mod.F90) -----------------------
// set_ltraj :375-533 keeps BL_DRIVER (the nonlinear scheme that produces the diagonals, once per trajectory time level) on the
// caller's side of the boundary; the decomposition, p^kappa and every TL / AD / NL application run here.
int fv3lm_turb_set_ltraj(fv3lm_handle* h, int slot, const fv3lm_turb_coeffs* co) {
  FV3LM_TRY_H(h)
  ensure_runner(h);
  StepRunner* r = h->step; const Geom& g = h->dv.g;
  if (!co) throw std::runtime_error("fv3lm_turb_set_ltraj: null coefficients");
  if (g.K < 2) throw std::runtime_error("fv3lm_turb_set_ltraj: needs npz >= 2");
  if (slot < 0) throw std::runtime_error("fv3lm: negative trajectory slot");
  if ((int)r->turb.size() <= slot) r->turb.resize(slot + 1);
  TurbLtraj& lt = r->turb[slot];
  const size_t n = compact_doubles(g, g.K);
  const double* src[TURB_NARR] = {co->akv, co->bkv, co->ckv, co->aks, co->bks, co->cks, co->akq, co->bkq, co->ckq, co->pk};
  static const char* nm[TURB_NARR] = {"akv", "bkv", "ckv", "aks", "bks", "cks", "akq", "bkq", "ckq", "pk"};
  for (int a = 0; a < TURB_NARR; a++) {
    if (!lt.d[a]) lt.d[a] = (double*)dev::alloc(n * sizeof(double));
    if (src[a]) dev::h2d(lt.d[a], src[a], n * sizeof(double));
    else if (a != TURB_PK) throw std::runtime_error(std::string("fv3lm_turb_set_ltraj: missing array ") + nm[a]);
  }
  const TurbDims s{g.ie, g.je, g.ntile, g.K};
  if (!co->pk) {
    if ((int)r->slots.size() <= slot || r->slots[slot].empty())
      throw std::runtime_error("fv3lm_turb_set_ltraj: pk not given and trajectory slot not set (delp is needed)");
    turb_pk(s, slot_field(h, slot, 3), lt.d[TURB_PK], h->cfg.ptop, h->cfg.kappa);
  }
  if (!co->decomposed) turb_lu(s, lt);
  lt.set = true;
  dev::sync();
  FV3LM_CATCH(h)
}

static void turb_apply(fv3lm_handle* h, int slot, double* const* f10, bool adjoint) {
  StepRunner* r = h->step; const Geom& g = h->dv.g;
  if (slot < 0 || (int)r->turb.size() <= slot || !r->turb[slot].set)
    throw std::runtime_error("fv3lm_turb_step: fv3lm_turb_set_ltraj was not called for this slot");
  double* f7[7] = {f10[0], f10[1], f10[2], f10[4], f10[6], f10[5], f10[7]};      // u v t qv qi ql o3 (order of step_tl :259-265)
  turb_solve(TurbDims{g.ie, g.je, g.ntile, g.K}, r->turb[slot], f7, pow(1.0e5, h->cfg.kappa), adjoint);
}

int fv3lm_turb_step_tl_dev(fv3lm_handle* h, int slot) {
  FV3LM_TRY_H(h)
  ensure_runner(h);
  turb_apply(h, slot, h->step->pert, false);
  FV3LM_CATCH(h)
}

int fv3lm_turb_step_ad_dev(fv3lm_handle* h, int slot) {
  FV3LM_TRY_H(h)
  ensure_runner(h);
  turb_apply(h, slot, h->step->pert, true);
  FV3LM_CATCH(h)
}

int fv3lm_turb_step_tl(fv3lm_handle* h, int slot, fv3lm_fields* pert) {
  int rc = fv3lm_pert_upload(h, pert);
  if (!rc) rc = fv3lm_turb_step_tl_dev(h, slot);
  if (!rc) rc = fv3lm_pert_download(h, pert);
  return rc;
}

int fv3lm_turb_step_ad(fv3lm_handle* h, int slot, fv3lm_fields* pert) {
  int rc = fv3lm_pert_upload(h, pert);
  if (!rc) rc = fv3lm_turb_step_ad_dev(h, slot);
  if (!rc) rc = fv3lm_pert_download(h, pert);
  return rc;
}

// step_nl :149-213: the same solves applied to the trajectory fields of slot_state, in place, with the local trajectory of
// slot_ltraj (fv3jedi_lm_mod calls the physics after the dynamics: the state has moved on, src/fv3jedi_lm_mod.F90:153-156)
int fv3lm_turb_step_nl(fv3lm_handle* h, int slot_ltraj, int slot_state) {
  FV3LM_TRY_H(h)
  ensure_runner(h);
  StepRunner* r = h->step;
  if (slot_state < 0 || (int)r->slots.size() <= slot_state || r->slots[slot_state].empty())
    throw std::runtime_error("fv3lm_turb_step_nl: trajectory slot not set");
  turb_apply(h, slot_ltraj, r->slots[slot_state].data(), false);
  dev::sync();
  FV3LM_CATCH(h)
}

// Timed loop on the library's own stream (CUDA events): `iters` repetitions of one TL step followed
// by one AD step on resident data.  ms[0] = TL ms/step, ms[1] = AD ms/step.  Used by bench.py.
int fv3lm_time_steps(fv3lm_handle* h, int slot, int warmup, int iters, double* ms) {
  FV3LM_TRY_H(h)
  ensure_runner(h);
#ifndef FV3LM_HOST_EMU
  cudaEvent_t e0, e1, e2;
  cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  for (int n = 0; n < warmup; n++) { run_step(h, slot, MODE_TL); run_step(h, slot, MODE_AD); }
  dev::sync();
  double tl = 0.0, ad = 0.0;
  for (int n = 0; n < iters; n++) {
    cudaEventRecord(e0, dev::stream());
    run_step(h, slot, MODE_TL);
    cudaEventRecord(e1, dev::stream());
    run_step(h, slot, MODE_AD);
    cudaEventRecord(e2, dev::stream());
    cudaEventSynchronize(e2);
    float a = 0, b = 0;
    cudaEventElapsedTime(&a, e0, e1); cudaEventElapsedTime(&b, e1, e2);
    tl += a; ad += b;
  }
  ms[0] = tl / iters; ms[1] = ad / iters;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
#else
  (void)slot; (void)warmup; (void)iters; ms[0] = ms[1] = 0.0;
  throw std::runtime_error("fv3lm_time_steps: host emulation build has no timer");
#endif
  FV3LM_CATCH(h)
}

// The same for the turbulence solves: ms[0] = TL ms/call, ms[1] = AD ms/call on the device-resident increments.
int fv3lm_time_turb(fv3lm_handle* h, int slot, int warmup, int iters, double* ms) {
  FV3LM_TRY_H(h)
  ensure_runner(h);
#ifndef FV3LM_HOST_EMU
  cudaEvent_t e0, e1, e2;
  cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  for (int n = 0; n < warmup; n++) { turb_apply(h, slot, h->step->pert, false); turb_apply(h, slot, h->step->pert, true); }
  dev::sync();
  double tl = 0.0, ad = 0.0;
  for (int n = 0; n < iters; n++) {
    cudaEventRecord(e0, dev::stream());
    turb_apply(h, slot, h->step->pert, false);
    cudaEventRecord(e1, dev::stream());
    turb_apply(h, slot, h->step->pert, true);
    cudaEventRecord(e2, dev::stream());
    cudaEventSynchronize(e2);
    float a = 0, b = 0;
    cudaEventElapsedTime(&a, e0, e1); cudaEventElapsedTime(&b, e1, e2);
    tl += a; ad += b;
  }
  ms[0] = tl / iters; ms[1] = ad / iters;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
#else
  (void)slot; (void)warmup; (void)iters; ms[0] = ms[1] = 0.0;
  throw std::runtime_error("fv3lm_time_turb: host emulation build has no timer");
#endif
  FV3LM_CATCH(h)
}

// Profiling pass: per-op CUDA-event times of `iters` TL+AD steps (serialised; not a bench number).
// Writes lines "name launches total_ms alg_bytes" into buf.
int fv3lm_profile_steps(fv3lm_handle* h, int slot, int iters, char* buf, int buflen) {
  FV3LM_TRY_H(h)
  ensure_runner(h);
  dev::prof.clear();
  struct ProfGuard { ProfGuard() { dev::profiling = true; } ~ProfGuard() { dev::profiling = false; } };   // also reset when a step throws
  {
    ProfGuard guard;
    for (int n = 0; n < iters; n++) { run_step(h, slot, MODE_TL); run_step(h, slot, MODE_AD); }
    dev::sync();
  }
  std::string out;
  for (auto& kv : dev::prof) {
    char line[512];
    snprintf(line, sizeof(line), "%s %lld %.6f %.0f\n", kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.alg_bytes);
    out += line;
  }
  if ((int)out.size() + 1 > buflen) out.resize(buflen > 0 ? buflen - 1 : 0);
  if (buflen > 0) { memcpy(buf, out.c_str(), out.size()); buf[out.size()] = 0; }
  FV3LM_CATCH(h)
}

// structural statistics of a module's program: ops, values, bytes if every value is kept (AD)
int fv3lm_program_stats(fv3lm_handle* h, const char* module, double* out) {
  FV3LM_TRY_H(h)
  ModuleParams prm; prm.cfg = &h->cfg; prm.ak = &h->ak; prm.bk = &h->bk; prm.v["bdt"] = h->cfg.dt;
  Program P; P.dv = &h->dv; ModuleIO io;
  build_module(module, P, h->mo, io, prm);
  double bytes = 0;
  for (size_t i = 0; i < P.vals.size(); i++) bytes += (double)P.val_doubles((int)i) * 8.0;
  int npatch = 0; for (auto& o : P.ops) npatch += o.inplace ? 1 : 0;
  out[0] = (double)P.ops.size(); out[1] = (double)P.vals.size(); out[2] = bytes; out[3] = npatch;
  FV3LM_CATCH(h)
}

}  // extern "C"
