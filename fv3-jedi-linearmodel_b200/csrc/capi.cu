// C ABI (include/fv3lm_b200.h): handle, metrics upload, module runner.
#include "capi_internal.h"

using namespace fv3lm;
namespace fv3lm { void a2b_corner_weights(const Geom& g, const double* glon, const double* glat, const double* alon, const double* alat, double* out); }

thread_local std::string fv3lm_g_err;
static_assert(sizeof(fv3lm_config) == 360, "fv3lm_config is mirrored field by field in fv3lm.py and fortran/fv3lm_b200_capi_mod.F90");

// host [rows][NX] contiguous <-> device [rows][pitch]
static void up2d(const Geom& g, double* d, const double* h, size_t rows) {
#ifndef FV3LM_HOST_EMU
  if (cudaMemcpy2DAsync(d, g.pitch * sizeof(double), h, g.NX * sizeof(double), g.NX * sizeof(double), rows,
                        cudaMemcpyHostToDevice, dev::stream()) != cudaSuccess)
    throw std::runtime_error("fv3lm: cudaMemcpy2D h2d failed");
#else
  for (size_t r = 0; r < rows; r++) memcpy(d + r * g.pitch, h + r * g.NX, g.NX * sizeof(double));
#endif
}
static void down2d(const Geom& g, double* h, const double* d, size_t rows) {
#ifndef FV3LM_HOST_EMU
  if (cudaMemcpy2DAsync(h, g.NX * sizeof(double), d, g.pitch * sizeof(double), g.NX * sizeof(double), rows,
                        cudaMemcpyDeviceToHost, dev::stream()) != cudaSuccess)
    throw std::runtime_error("fv3lm: cudaMemcpy2D d2h failed");
#else
  for (size_t r = 0; r < rows; r++) memcpy(h + r * g.NX, d + r * g.pitch, g.NX * sizeof(double));
#endif
}

extern "C" {

int fv3lm_create(const fv3lm_config* cfg, const double* ak, const double* bk, fv3lm_handle** out) {
  fv3lm_handle* h = nullptr;
  FV3LM_TRY
  if (!cfg || !out) throw std::runtime_error("fv3lm_create: null argument");
  if (cfg->ntiles != 6 || cfg->npx != cfg->npy || cfg->ng != 3) throw std::runtime_error("fv3lm_create: need 6 square tiles and ng = 3");
  if (cfg->npz > 95 || cfg->npz < 1) throw std::runtime_error("fv3lm_create: npz must be in 1..95");
  for (int ho : {cfg->hord_mt, cfg->hord_vt, cfg->hord_tm, cfg->hord_dp, cfg->hord_tr})
    if (ho != 1 && ho != 2 && ho != 333) throw std::runtime_error("fv3lm_create: hord_* must be 1, 2 or 333 (the linear schemes the TL/AD implement, tp_core_tlm.F90:2431-2488)");
  if (cfg->two_sided)
    for (int ho : {cfg->traj.hord_mt, cfg->traj.hord_vt, cfg->traj.hord_tm, cfg->traj.hord_dp, cfg->traj.hord_tr})
      if (ho != 333 && !(ho >= 1 && ho <= 13))
        throw std::runtime_error("fv3lm_create: traj.hord_* must be in 1..13 or 333");
  if (cfg->two_sided)
    for (int ko : {cfg->traj.kord_mt, cfg->traj.kord_wz, cfg->traj.kord_tm, cfg->traj.kord_tr}) {
      const int a = ko < 0 ? -ko : ko;
      if (!(a == 0 || a == 17 || (a >= 8 && a <= 14))) throw std::runtime_error("fv3lm_create: traj.kord_* must be 8..14 (monotone profiles of the nonlinear model), 17 or 0 (linear)");
    }
  if (cfg->nq != 4) throw std::runtime_error("fv3lm_create: nq must be 4 (qv, ql, qi, o3)");
  if (cfg->n_split < 1 || cfg->k_split < 1 || !(cfg->dt > 0.0)) throw std::runtime_error("fv3lm_create: n_split, k_split and dt must be positive");
  if (cfg->nord < 0 || cfg->nord > 3) throw std::runtime_error("fv3lm_create: nord must be in 0..3");
  if (cfg->q_split_dynamic != 0 && cfg->q_split_dynamic != 1) throw std::runtime_error("fv3lm_create: q_split_dynamic must be 0 or 1");
  if (cfg->q_split_max < 0 || cfg->q_split_max > 8)
    throw std::runtime_error("fv3lm_create: q_split_max must be in 0..8 (0 = 3; every issued tracer sub-step is part of the static program)");
  if (!cfg->hydrostatic && cfg->a_imp != 0.0 && !(cfg->a_imp > 0.5))
    throw std::runtime_error("fv3lm_create: a_imp <= 0.5 selects the RIM_2D / SIM3 solvers (model/nh_core_nlm.F90:136-146), which are not built; use 0.5 < a_imp <= 1");
  if (cfg->npx < 9) throw std::runtime_error("fv3lm_create: need at least 8 cells per tile edge");
  // switches whose paths are not built change the dynamics in the reference: refuse them rather than ignore them
  if (cfg->beta < 0.0 || cfg->beta >= 1.0)
    throw std::runtime_error("fv3lm_create: beta must be in [0, 1): beta > 0 selects grad1_p_update / split_p_grad, beta < -0.1 the one_grad_p branch of the non-hydrostatic core (model/dyn_core_nlm.F90:865-877), which is not built");
  if (cfg->d_ext < 0.0) throw std::runtime_error("fv3lm_create: d_ext < 0");
  // d_ext > 0: external-mode divergence damping of the hydrostatic core (model/dyn_core_nlm.F90:642-726, one_grad_p :1713-1727).  The
  // non-hydrostatic pressure gradients (nh_p_grad, split_p_grad) never read divg2, so there the switch changes nothing (checked against the
  // reference's own DYN_CORE_TLM: bit-identical outputs with d_ext = 0.02 and 0) and is accepted as is.
  int device = -1;
#ifndef FV3LM_HOST_EMU
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    throw std::runtime_error("fv3lm_create: no CUDA device -- this library has no CPU fallback");
  // bind this process to its GPU before anything is allocated (several MPI ranks per node would otherwise all land on device 0)
  if (cfg->device > ndev) throw std::runtime_error("fv3lm_create: device index out of range");
  if (cfg->device != 0) {
    device = cfg->device > 0 ? cfg->device - 1 : (cfg->rank < 0 ? 0 : cfg->rank) % ndev;
    if (cudaSetDevice(device) != cudaSuccess) throw std::runtime_error("fv3lm_create: cudaSetDevice failed");
  } else if (cudaGetDevice(&device) != cudaSuccess) throw std::runtime_error("fv3lm_create: cudaGetDevice failed");
#endif
  h = new fv3lm_handle();
  h->cfg = *cfg;
  h->device = device;
  Geom& g = h->dv.g;
  g.N = cfg->npx - 1; g.npx = cfg->npx; g.npy = cfg->npy; g.ng = cfg->ng;
  Decomp& dc = h->dc;
  dc.init(g.N, cfg->rank, cfg->nranks < 1 ? 1 : cfg->nranks, cfg->layout_x, cfg->layout_y);
  if (dc.per_rank > MAXSUB) throw std::runtime_error("fv3lm_create: too many sub-domains per rank");
  g.is = 1; g.ie = dc.nxl; g.js = 1; g.je = dc.nyl;
  g.NX = dc.nxl + 2 * g.ng + 1; g.NY = dc.nyl + 2 * g.ng + 1;
  g.pitch = (g.NX + 3) / 4 * 4;
  g.ntile = dc.per_rank; g.K = cfg->npz; g.slab = g.pitch * g.NY;
  if ((double)g.ntile * (g.K + 1) * g.slab >= 2147483647.0) throw std::runtime_error("fv3lm_create: a field on this rank would exceed 2^31 elements (kernels use 32-bit offsets): use more ranks");
  if (g.slab < g.N + 2 * g.ng + 1) throw std::runtime_error("fv3lm_create: sub-domain slab smaller than a tile edge");
  for (int l = 0; l < MAXSUB; l++) { g.i0[l] = 0; g.j0[l] = 0; g.tile_of[l] = 0; }
  for (int l = 0; l < dc.per_rank; l++) {
    const int d = dc.global_id(dc.rank, l);
    g.i0[l] = (short)dc.i0_of(d); g.j0[l] = (short)dc.j0_of(d); g.tile_of[l] = (short)dc.tile_of(d);
  }
  h->comm.rank = dc.rank; h->comm.nranks = dc.nranks;
  h->dv.comm = &h->comm;
  if (ak && bk) { h->ak.assign(ak, ak + cfg->npz + 1); h->bk.assign(bk, bk + cfg->npz + 1); }
  memset(&h->dv.m, 0, sizeof(Metrics));
  h->mo.build(g, dc, &h->comm);
  *out = h;
  FV3LM_CATCH(h)
}

int fv3lm_destroy(fv3lm_handle* h) {
  if (!h) return 0;
  FV3LM_TRY_H(h)
  dev::sync();
  for (auto& kv : h->metric_dev) dev::free_(kv.second);
  if (h->step) {
    StepRunner* r = h->step;
    for (auto& kv : r->io.inputs) { dev::free_(r->P.vals[kv.second].traj); dev::free_(r->P.vals[kv.second].pert); }
    for (auto& kv : r->io.outputs) if (!r->io.is_input(kv.second)) { dev::free_(r->P.vals[kv.second].traj); dev::free_(r->P.vals[kv.second].pert); }
    for (int f = 0; f < NFIELD; f++) dev::free_(r->pert[f]);
    dev::free_(r->phis);
    for (auto& s : r->slots) for (double* p : s) dev::free_(p);
    for (auto& lt : r->turb) for (double* p : lt.d) dev::free_(p);
    if (r->c2l) {
      for (const char* n : {"a11", "a12", "a21", "a22", "ua", "va"}) dev::free_(r->c2l->P.vals[r->c2l->id[n]].traj);
      dev::free_(r->c2l->ua); dev::free_(r->c2l->va);
      delete r->c2l;
    }
#ifndef FV3LM_HOST_EMU
    for (auto& sg : r->graph) if (sg.exec) cudaGraphExecDestroy((cudaGraphExec_t)sg.exec);
#endif
    delete r;
  }
  h->mo.destroy();
  h->comm.destroy();
  delete h;
  h = nullptr;
  FV3LM_CATCH(h)
}

const char* fv3lm_last_error(const fv3lm_handle* h) { return h ? h->err.c_str() : fv3lm_g_err.c_str(); }

int fv3lm_set_metric(fv3lm_handle* h, const char* name, const double* host, int is_1d) {
  FV3LM_TRY_H(h)
  const Geom& g = h->dv.g;
  std::string nm(name);
  double*& d = h->metric_dev[nm];
  if (!d) { d = (double*)dev::alloc((size_t)g.ntile * g.slab * sizeof(double)); dev::zero(d, (size_t)g.ntile * g.slab * sizeof(double)); }
  if (is_1d) {
    const int nxg = g.N + 2 * g.ng + 1;   // 1-D edge factors are whole-tile arrays (gridstruct%edge_w(npy) ...)
    for (int t = 0; t < g.ntile; t++) dev::h2d(d + (size_t)t * g.slab, host + (size_t)t * nxg, nxg * sizeof(double));
  } else {
    up2d(g, d, host, (size_t)g.ntile * g.NY);
  }
  dev::sync();
  bool found = false;
#define X(n) if (nm == #n) { h->dv.m.n = d; found = true; }
  FV3LM_METRIC_LIST(X)
#undef X
  if (!found) throw std::runtime_error("fv3lm_set_metric: unknown metric " + nm);
  if (nm == "dxa" || nm == "dya") {
    // weights of q(e-2 .. e+1) in the PPM edge value AL(e) on a cube edge (e = 1 or np along the sweep direction): the mean of the two
    // one-sided extrapolations of tp_core_tlm.F90:2405-2417 (x) / :2573-2585 (y); S_ppm::al_w's formulas, evaluated once per handle
    const bool xdir = nm == "dxa";
    const int lo = g.ng - 1;
    std::vector<double> w[4];
    for (int n = 0; n < 4; n++) w[n].assign((size_t)g.ntile * g.NY * g.NX, 0.0);
    for (int t = 0; t < g.ntile; t++) {
      const int cpos = (xdir ? g.i0[t] : g.j0[t]) - lo, np = xdir ? g.npx : g.npy;
      for (int jj = 0; jj < g.NY; jj++)
        for (int ii = 0; ii < g.NX; ii++) {
          const int a = xdir ? ii : jj, e = a + cpos, len = xdir ? g.NX : g.NY;
          if ((e != 1 && e != np) || a - 2 < 0 || a + 1 >= len) continue;
          auto D = [&](int d) { return host[((size_t)t * g.NY + (xdir ? jj : jj + d)) * g.NX + (xdir ? ii + d : ii)]; };
          const double a0 = D(-1), am = D(-2), a1 = D(0), a2 = D(1);
          const size_t o = ((size_t)t * g.NY + jj) * g.NX + ii;
          w[0][o] = -0.5 * a0 / (am + a0); w[1][o] = 0.5 * (2.0 * a0 + am) / (am + a0);
          w[2][o] = 0.5 * (2.0 * a1 + a2) / (a1 + a2); w[3][o] = -0.5 * a1 / (a1 + a2);
        }
    }
    for (int n = 0; n < 4; n++) {
      const std::string wn = std::string(xdir ? "ppmw_x" : "ppmw_y") + char('0' + n);
      double*& dw = h->metric_dev[wn];
      if (!dw) { dw = (double*)dev::alloc((size_t)g.ntile * g.slab * sizeof(double)); dev::zero(dw, (size_t)g.ntile * g.slab * sizeof(double)); }
      up2d(g, dw, w[n].data(), (size_t)g.ntile * g.NY);
      double** slot = nullptr;
      if (xdir) slot = n == 0 ? (double**)&h->dv.m.ppmw_x0 : n == 1 ? (double**)&h->dv.m.ppmw_x1 : n == 2 ? (double**)&h->dv.m.ppmw_x2 : (double**)&h->dv.m.ppmw_x3;
      else slot = n == 0 ? (double**)&h->dv.m.ppmw_y0 : n == 1 ? (double**)&h->dv.m.ppmw_y1 : n == 2 ? (double**)&h->dv.m.ppmw_y2 : (double**)&h->dv.m.ppmw_y3;
      *slot = dw;
    }
    dev::sync();
  }
  if (nm == "grid_lon" || nm == "grid_lat" || nm == "agrid_lon" || nm == "agrid_lat") {
    h->geo_host[nm].assign(host, host + (size_t)g.ntile * g.NY * g.NX);
    if (h->geo_host.size() == 4) {
      // a2b_ord4 corner extrapolation weights (a2b_edge_nlm.F90:73-106): geometry only, computed once
      std::vector<double> w((size_t)g.ntile * 12), slabs((size_t)g.ntile * g.slab, 0.0);
      a2b_corner_weights(g, h->geo_host["grid_lon"].data(), h->geo_host["grid_lat"].data(), h->geo_host["agrid_lon"].data(),
                         h->geo_host["agrid_lat"].data(), w.data());
      for (int t = 0; t < g.ntile; t++) for (int n = 0; n < 12; n++) slabs[(size_t)t * g.slab + n] = w[t * 12 + n];
      double*& dw = h->metric_dev["a2b_cw"];
      if (!dw) dw = (double*)dev::alloc(slabs.size() * sizeof(double));
      dev::h2d(dw, slabs.data(), slabs.size() * sizeof(double));
      dev::sync();
      h->dv.m.a2b_cw = dw;
    }
  }
  FV3LM_CATCH(h)
}

int fv3lm_set_metric_scalar(fv3lm_handle* h, const char* name, double value) {
  FV3LM_TRY_H(h)
  std::string nm(name);
  if (nm == "da_min") h->dv.m.da_min = value;
  else if (nm == "da_min_c") h->dv.m.da_min_c = value;
  else throw std::runtime_error("fv3lm_set_metric_scalar: unknown " + nm);
  FV3LM_CATCH(h)
}

const char* fv3lm_module_list(void) { return module_list(); }

int fv3lm_module_run(fv3lm_handle* h, const char* module, int mode, int nfields, const char* const* names,
                     double* const* traj, double* const* pert, int nparams, const char* const* pnames,
                     const double* pvals) {
  FV3LM_TRY_H(h)
  const Geom& g = h->dv.g;
  ModuleParams prm;
  for (int n = 0; n < nparams; n++) prm.v[pnames[n]] = pvals[n];
  prm.cfg = &h->cfg; prm.ak = &h->ak; prm.bk = &h->bk;
  Program P; P.dv = &h->dv; P.name = module;
  ModuleIO io;
  build_module(module, P, h->mo, io, prm);
  std::map<std::string, int> idx;
  for (int n = 0; n < nfields; n++) idx[names[n]] = n;
  // bind externals
  std::vector<int> ext;
  for (auto& kv : io.inputs) ext.push_back(kv.second);
  for (auto& kv : io.outputs) if (!io.is_input(kv.second)) ext.push_back(kv.second);
  auto host_of = [&](int id, bool want_pert) -> double* {
    auto it = idx.find(P.vals[id].name);
    if (it == idx.end()) return nullptr;
    return want_pert ? pert[it->second] : traj[it->second];
  };
  for (auto& kv : io.inputs) {
    int id = kv.second; Value& v = P.vals[id];
    double* ht = host_of(id, false);
    if (!ht) throw std::runtime_error(std::string("module ") + module + ": missing input field " + v.name);
    v.traj = h->dv.pool.get(P.val_doubles(id));
    up2d(g, v.traj, ht, (size_t)g.ntile * v.nk * g.NY);
    v.active = (mode != MODE_NL) && host_of(id, true) != nullptr;
  }
  P.sweep_kind = (mode == MODE_NL || mode == MODE_TL) ? VAR_FWD : VAR_AD;
  P.analyse();
  for (int id : ext) {
    Value& v = P.vals[id];
    if (!v.traj) { v.traj = h->dv.pool.get(P.val_doubles(id)); dev::zero(v.traj, P.val_doubles(id) * sizeof(double)); }
    if (v.active) {
      v.pert = h->dv.pool.get(P.val_doubles(id));
      dev::zero(v.pert, P.val_doubles(id) * sizeof(double));
    }
  }
  if (mode == MODE_TL) {
    for (auto& kv : io.inputs) { Value& v = P.vals[kv.second]; if (v.active) up2d(g, v.pert, host_of(kv.second, true), (size_t)g.ntile * v.nk * g.NY); }
  } else if (mode == MODE_AD) {
    for (auto& kv : io.outputs) {
      Value& v = P.vals[kv.second];
      double* hp = host_of(kv.second, true);
      if (v.active && hp) up2d(g, v.pert, hp, (size_t)g.ntile * v.nk * g.NY);
    }
  }
  auto give_back = [&]() { for (int id : ext) { Value& v = P.vals[id]; h->dv.pool.put(v.traj); h->dv.pool.put(v.pert); v.traj = v.pert = nullptr; } };
  try { P.run((Mode)mode); }
  catch (...) { give_back(); throw; }       // (Program::run has already released its own intermediates)
  // download
  for (auto& kv : io.outputs) {
    Value& v = P.vals[kv.second];
    double* ht = host_of(kv.second, false);
    if (ht) down2d(g, ht, v.traj, (size_t)g.ntile * v.nk * g.NY);
    if (mode == MODE_TL) { double* hp = host_of(kv.second, true); if (hp && v.active) down2d(g, hp, v.pert, (size_t)g.ntile * v.nk * g.NY); }
  }
  if (mode == MODE_AD) {
    for (auto& kv : io.inputs) {
      Value& v = P.vals[kv.second];
      double* hp = host_of(kv.second, true);
      if (hp && v.active) down2d(g, hp, v.pert, (size_t)g.ntile * v.nk * g.NY);
    }
  }
  dev::sync();
  give_back();
  P.check_status_flags();
  FV3LM_CATCH(h)
}

// ---- domain decomposition / communicator ----------------------------------------------------
int fv3lm_decomp_info(const fv3lm_handle* h, int* out /* nsub, nxl, nyl, layout_x, layout_y, nsub_total */, int* tile, int* i0, int* j0) {
  if (!h) return 1;
  const Decomp& dc = h->dc;
  if (out) { out[0] = dc.per_rank; out[1] = dc.nxl; out[2] = dc.nyl; out[3] = dc.lx; out[4] = dc.ly; out[5] = dc.nsub_total; }
  for (int l = 0; l < dc.per_rank; l++) {
    const int d = dc.global_id(dc.rank, l);
    if (tile) tile[l] = dc.tile_of(d);
    if (i0) i0[l] = dc.i0_of(d);
    if (j0) j0[l] = dc.j0_of(d);
  }
  return 0;
}
int fv3lm_nccl_unique_id(char* out128) {
  fv3lm_handle* h = nullptr;
  FV3LM_TRY
  Comm::nccl_unique_id(out128);
  FV3LM_CATCH(h)
}
int fv3lm_comm_init_nccl(fv3lm_handle* h, const char* id128) {
  FV3LM_TRY_H(h)
  h->comm.init_nccl(id128);
  FV3LM_CATCH(h)
}
int fv3lm_comm_set_callback(fv3lm_handle* h, fv3lm_exchange_fn fn, void* user) {
  FV3LM_TRY_H(h)
  h->comm.cb = (ExchangeCallback)fn; h->comm.cb_user = user;
  FV3LM_CATCH(h)
}
int fv3lm_comm_stats(const fv3lm_handle* h, double* out2) {
  if (!h) return 1;
  out2[0] = (double)h->comm.n_exchanges; out2[1] = h->comm.bytes_sent;
  return 0;
}

long long fv3lm_launch_count(void) { return dev::launches; }
double fv3lm_pool_peak_bytes(const fv3lm_handle* h) { return h ? (double)h->dv.pool.bytes_peak : 0.0; }
int fv3lm_sync(fv3lm_handle* h) {
  FV3LM_TRY_H(h)
  dev::sync();
  FV3LM_CATCH(h)
}

}  // extern "C"
