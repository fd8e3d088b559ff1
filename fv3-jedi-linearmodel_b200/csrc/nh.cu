// Non-hydrostatic program builders: update_dz_c / update_dz_d, Riem solvers, and the
// non-hydrostatic branch of dyn_core (model/dyn_core_nlm.F90:78-1040).
#include "nh.h"
#include "stages_riem.h"
#include "modules.h"

namespace fv3lm {

static LevD dp_ref_of(const std::vector<double>& ak, const std::vector<double>& bk, int K) {
  LevD d; for (int k = 0; k < 96; k++) d.v[k] = 0.0;
  for (int k = 0; k < K; k++) d.v[k] = (ak[k + 1] - ak[k]) + (bk[k + 1] - bk[k]) * 1.e5;   // dyn_core_nlm.F90:213-215
  return d;
}

// Riem_Solver_c / Riem_Solver3 as level-parallel stages + thin column recurrences (stages_riem.h).
// FV3LM_RIEM_MONO=1 selects the monolithic one-thread-per-column kernel S_riem instead (kept for A/B measurements).
static bool riem_mono() { static const bool m = getenv("FV3LM_RIEM_MONO") != nullptr; return m; }

RiemOut build_riem(Program& P, const RiemPrm& r, int delp, int pt, int z, int w, int ws, int zb, const std::string& tag) {
  const int K = r.K, h = r.halo;
  auto nm = [&](const char* s) { return tag + "." + s; };
  if (!(r.a_imp > 0.5)) throw std::runtime_error("riem: a_imp <= 0.5 (RIM_2D / SIM3 solvers) is not supported");
  // Riem_Solver3 with 0.5 < a_imp <= 0.999: SIM_solver, off-centring alpha = a_imp (scale_m = 0); everything else: SIM1_solver
  const bool sim = r.mode == 1 && !(r.a_imp > 0.999);
  if (riem_mono() && !sim) {   // (the monolithic A/B kernel implements SIM1 only)
    int o0 = P.val(nm("pp"), K + 1), o1 = P.val(nm("z"), K + 1), o2 = P.val(nm("w"), K), o3 = P.val(nm("dz"), K);
    add_col<S_riem>(P, r.mode == 0 ? "riem_solver_c" : "riem_solver3", {K, r.mode, h, r.dt, r.akap, r.ptop, r.rdgas, r.grav, r.p_fac}, {delp, pt, z, w, ws, zb}, {o0, o1, o2, o3});
    return {o0, o1, o2, o3};
  }
  const double alpha = sim ? r.a_imp : 1.0, beta = 1.0 - alpha, ra = 1.0 / alpha, t2 = beta / alpha;
  const double gama = 1.0 / (1.0 - r.akap), rgrav = 1.0 / r.grav, t1g = sim ? 2.0 * gama * (alpha * r.dt) * (alpha * r.dt) : gama * 2.0 * r.dt * r.dt,
               rdt = 1.0 / r.dt, capa1 = r.akap - 1.0;
  int pem = P.val(nm("pem"), K + 1);
  add_col<S_cum>(P, "rs_pem", {K, 0, h, r.ptop, 1.0}, {delp, delp}, {pem});
  int pm2 = P.val(nm("pm2"), K), pe = P.val(nm("pe"), K);
  P.add<S_rs_pe>("rs_pe", {r.mode, h, gama, rgrav, r.rdgas}, {delp, pt, z, pem}, {pm2, pe}, K);
  int bb = P.val(nm("bb"), K), gg = P.val(nm("g"), K), dd = P.val(nm("dd"), K);
  P.add<S_rs_cpp>("rs_cpp", {K, h}, {delp, pe}, {bb, gg, dd}, K);
  int pp = P.val(nm("ppi"), K + 1);
  add_col<S_tri>(P, "rs_tri_pp", {K, 1, 0, 0, 1, h}, {bb, bb, gg, dd}, {pp});
  int A = P.val(nm("aa"), K + 1);
  P.add<S_rs_aa>("rs_aa", {K, h, t1g}, {z, pem, pp}, {A}, K + 1);
  int di = P.val(nm("di"), K), rhs = P.val(nm("rhs"), K);
  if (sim) P.add<S_rs_rw_t<true>>("rs_rw_sim", {K, h, rgrav, r.dt, t2, ra}, {delp, w, pp, A, ws}, {di, rhs}, K);
  else P.add<S_rs_rw>("rs_rw", {K, h, rgrav, r.dt, 0.0, 1.0}, {delp, w, pp, A, ws}, {di, rhs}, K);
  int w2 = P.val(nm("w2"), K);
  add_col<S_tri>(P, "rs_tri_w", {K, 0, 0, 1, 0, h}, {A, di, A, rhs}, {w2});
  int pe2 = P.val(nm("pe2"), K + 1);
  if (sim) add_col<S_rs_pe2_t<true>>(P, "rs_pe2_sim", {K, h, rgrav, rdt, beta, ra}, {delp, w2, w, pp}, {pe2});
  else add_col<S_rs_pe2>(P, "rs_pe2", {K, h, rgrav, rdt, 0.0, 1.0}, {delp, w2, w}, {pe2});
  int p1 = P.val(nm("p1"), K);
  add_col<S_rs_p1>(P, "rs_p1", {K, h}, {pe2, bb, gg}, {p1});
  int dzn = P.val(nm("dzn"), K);
  P.add<S_rs_dz>("rs_dz", {h, rgrav, r.rdgas, capa1, r.p_fac}, {delp, pt, pm2, p1}, {dzn}, K);
  int zn = P.val(nm("zn"), K + 1);
  add_col<S_cum>(P, "rs_z", {K, 1, h, 0.0, r.mode == 0 ? r.grav : 1.0}, {dzn, zb}, {zn});
  if (r.mode == 0) {
    int pef = P.val(nm("pef"), K + 1);
    P.add<S_rs_pef>("rs_pef", {h, r.ptop}, {pe2, pem}, {pef}, K + 1);
    return {pef, zn, -1, -1};
  }
  if (sim) {
    int ppe = P.val(nm("ppe"), K + 1);
    P.add<S_rs_blend>("rs_blend", {h, beta}, {pe2, pp}, {ppe}, K + 1);
    return {ppe, zn, w2, dzn};
  }
  return {pe2, zn, w2, dzn};
}

// update_dz_c (model/nh_utils_nlm.F90:43-182).  gz is patched in place (fill_4corners); returns {gz_new, ws}
std::pair<int, int> build_update_dz_c(Program& P, Mosaic& mo, const LevD& dp0, double dt, int zs, int ut, int vt, int gz, const std::string& tag) {
  const int K = P.dv->g.K;
  auto nm = [&](const char* s) { return tag + "." + s; };
  int xfx = P.val(nm("xfx"), K + 1), yfx = P.val(nm("yfx"), K + 1), fx = P.val(nm("fx"), K + 1), fy = P.val(nm("fy"), K + 1);
  P.add<S_dzc_wind>("dzc_wind", {dp0, K}, {ut, vt}, {xfx, yfx}, K + 1);
  add_patch(P, "fill4c_x.gz", &mo.f4c1, {gz});
  P.add<S_dzc_flux<0>>("dzc_fx", {0}, {xfx, gz}, {fx}, K + 1);
  add_patch(P, "fill4c_y.gz", &mo.f4c2, {gz});
  P.add<S_dzc_flux<1>>("dzc_fy", {0}, {yfx, gz}, {fy}, K + 1);
  int gzn = P.val(nm("gzn"), K + 1), gzo = P.val(nm("gz"), K + 1), ws = P.val(nm("ws"), 1);
  P.add<S_dzc_upd>("dzc_upd", {0}, {gz, fx, fy, xfx, yfx}, {gzn}, K + 1);
  add_col<S_dz_clamp>(P, "dzc_clamp", {K, 1.0 / dt, 1}, {gzn, zs}, {gzo, ws});
  return {gzo, ws};
}

// update_dz_d (model/nh_utils_nlm.F90:183-296).  zh is patched in place (copy_corners); returns {zh_new, ws}
std::pair<int, int> build_update_dz_d(Program& P, Mosaic& mo, const LevD& dp0, const DswParams& dp, int hord_tm, double rdt, int zs, int zh,
                                      int crx, int cry, int xfx, int yfx, const std::string& tag, int hord_tm_pert) {
  const Geom& g = P.dv->g;
  const int K = g.K, is = g.is, ie = g.ie, js = g.js, je = g.je, ng = g.ng;
  auto nm = [&](const char* s) { return tag + "." + s; };
  int crxa = P.val(nm("crx_adv"), K + 1), xfxa = P.val(nm("xfx_adv"), K + 1), crya = P.val(nm("cry_adv"), K + 1), yfxa = P.val(nm("yfx_adv"), K + 1);
  add_col<S_edge_profile>(P, "edge_profile_crx", S_edge_profile::make(dp0, K, is, ie + 1, js - ng, je + ng), {crx}, {crxa});
  add_col<S_edge_profile>(P, "edge_profile_xfx", S_edge_profile::make(dp0, K, is, ie + 1, js - ng, je + ng), {xfx}, {xfxa});
  add_col<S_edge_profile>(P, "edge_profile_cry", S_edge_profile::make(dp0, K, is - ng, ie + ng, js, je + 1), {cry}, {crya});
  add_col<S_edge_profile>(P, "edge_profile_yfx", S_edge_profile::make(dp0, K, is - ng, ie + ng, js, je + 1), {yfx}, {yfxa});
  int ra_x = P.val(nm("ra_x"), K + 1), ra_y = P.val(nm("ra_y"), K + 1);
  P.add<S_ra>("dzd_ra", {0}, {xfxa, yfxa}, {ra_x, ra_y}, K + 1);
  LevOrd ho; for (int k = 0; k < 128; k++) ho.v[k] = (signed char)enc_hord(hord_tm);
  TpOut f;
  if (hord_tm_pert == 0 || hord_tm_pert == hord_tm) f = build_fv_tp_2d(P, mo, zh, crxa, crya, xfxa, yfxa, ra_x, ra_y, -1, -1, ho, K + 1, tag + ".tp_zh");
  else {   // model_tlmadm/nh_utils_tlm.F90:496-560: perturbation with hord_tm_pert, trajectory with hord_tm
    LevOrd hp; for (int k = 0; k < 128; k++) hp.v[k] = (signed char)enc_hord(hord_tm_pert);
    P.tl_only = true;
    TpOut a = build_fv_tp_2d(P, mo, zh, crxa, crya, xfxa, yfxa, ra_x, ra_y, -1, -1, hp, K + 1, tag + ".tp_zh_p");
    P.tl_only = false;
    TpOut b = build_fv_tp_2d(P, mo, P.detached(zh), P.detached(crxa), P.detached(crya), P.detached(xfxa), P.detached(yfxa), P.detached(ra_x), P.detached(ra_y),
                             -1, -1, ho, K + 1, tag + ".tp_zh_t");
    f.fx = P.val(tag + ".fx", K + 1); f.fy = P.val(tag + ".fy", K + 1);
    P.add<S_splice>("splice", {0}, {a.fx, b.fx}, {f.fx}, K + 1);
    P.add<S_splice>("splice", {0}, {a.fy, b.fy}, {f.fy}, K + 1);
  }
  // del6_vt_flux(ndif(k), damp(k)) with ndif = nord_v, damp = damp_vt; level K+1 repeats level K (:220-221)
  LevOrd nv; LevD dm, on; bool any = false;
  for (int k = 0; k < 128; k++) nv.v[k] = -1;
  for (int k = 0; k < 96; k++) { dm.v[k] = 0.0; on.v[k] = 0.0; }
  for (int k = 0; k <= K; k++) {
    const int ks = k < K ? k : K - 1;
    if (dp.damp_v.v[ks] > 1.e-5) { nv.v[k] = dp.nord_v.v[ks]; dm.v[k] = dp.damp_v.v[ks]; on.v[k] = 1.0; any = true; }
  }
  int fx2 = f.fx, fy2 = f.fy;
  if (any) { auto d = build_deln_public(P, mo, zh, nv, dm, K + 1, tag + ".del6"); fx2 = d.first; fy2 = d.second; }
  int zhn = P.val(nm("zhn"), K + 1), zho = P.val(nm("zh"), K + 1), ws = P.val(nm("ws"), 1);
  P.add<S_dzd_upd>("dzd_upd", {on}, {zh, f.fx, f.fy, ra_x, ra_y, fx2, fy2}, {zhn}, K + 1);
  add_col<S_dz_clamp>(P, "dzd_clamp", {K, rdt, 0}, {zhn, zs}, {zho, ws});
  return {zho, ws};
}

DynOut build_dyn_core_nh(Program& P, Mosaic& mo, const DynConfig& c, const std::vector<double>& ak, const std::vector<double>& bk,
                         DynState s, const std::string& tag) {
  const Geom& g = P.dv->g;
  const int K = g.K, is = g.is, ie = g.ie, js = g.js, je = g.je, ng = g.ng;
  const int isd = is - ng, ied = ie + ng, jsd = js - ng, jed = je + ng;
  const double dt = c.bdt / c.n_split, dt2 = 0.5 * dt, rdt = 1.0 / dt;
  DswParams dp; level_params(c, K, dp); dp.dt = dt; dp.hydrostatic = false;
  DswParams dpp; const bool two = level_params_pert(c, K, dpp); dpp.dt = dt; dpp.hydrostatic = false;
  const LevD dp0 = dp_ref_of(ak, bk, K);
  DynOut o;
  int u = s.u, v = s.v, pt = s.pt, delp = s.delp, w = s.w, delz = s.delz;
  int mfx = -1, mfy = -1, cx = -1, cy = -1, heat = -1, du_prev = -1, dv_prev = -1;
  int zs = P.val(tag + ".zs", 1);
  P.add<S_scale>("zs", {1.0 / c.grav, ng}, {s.phis}, {zs}, 1);
  int zh = -1, ws_d = -1;
  for (int it = 1; it <= c.n_split; it++) {
    const std::string tg = tag + ".it" + std::to_string(it);
    P.mark_segment();
    add_patch(P, "halo_w", &mo.h_center, {w});
    int gz;
    if (it == 1) {
      gz = P.val(tg + ".gz0", K + 1);
      add_col<S_gz_init>(P, "gz_init", {K}, {delz, zs}, {gz});
      add_patch(P, "halo_gz", &mo.h_center, {gz});
      zh = gz;
    }
    // gz for the C-grid half step is a copy of zh (update_dz_c patches and replaces it)
    gz = P.val(tg + ".gzc", K + 1);
    P.add<S_scale>("gz_copy", {1.0, ng}, {zh}, {gz}, K + 1);
    CswOut cs = build_c_sw(P, mo, delp, pt, u, v, w, dt2, false, c.nord, K, tg + ".csw");
    if (c.nord > 0) add_patch(P, "halo_divgd", &mo.h_corner, {cs.divg_d});
    auto dzc = build_update_dz_c(P, mo, dp0, dt2, zs, cs.ut, cs.vt, gz, tg + ".dzc");
    RiemOut rc = build_riem(P, {K, 0, 1, dt2, c.akap, c.ptop, c.rdgas, c.grav, c.p_fac, c.a_imp}, cs.delpc, cs.ptc, dzc.first, cs.wc, dzc.second, s.phis, tg + ".rsc");
    const int pef = rc.pp, gzr = rc.z;
    int uc = P.val(tg + ".uc", K), vc = P.val(tg + ".vc", K);
    P.add<S_pgrad_c>("p_grad_c", {dt2, 0}, {cs.uc, cs.vc, pef, gzr, cs.delpc}, {uc, vc}, K);
    add_patch(P, "halo_ucvc", &mo.h_cgrid, {uc, vc});
    DswOut ds = build_d_sw(P, mo, delp, pt, u, v, w, uc, vc, cs.ua, cs.va, cs.divg_d, dp, K, tg + ".dsw", two ? &dpp : nullptr);
    if (mfx < 0) { mfx = ds.fx; mfy = ds.fy; cx = ds.crx; cy = ds.cry; }
    else {
      int a = P.val(tg + ".mfx", K), b = P.val(tg + ".mfy", K), cc = P.val(tg + ".cx", K), d = P.val(tg + ".cy", K);
      P.add<S_add2>("acc_mfx", {is, ie + 1, js, je}, {mfx, ds.fx}, {a}, K);
      P.add<S_add2>("acc_mfy", {is, ie, js, je + 1}, {mfy, ds.fy}, {b}, K);
      P.add<S_add2>("acc_cx", {is, ie + 1, jsd, jed}, {cx, ds.crx}, {cc}, K);
      P.add<S_add2>("acc_cy", {isd, ied, js, je + 1}, {cy, ds.cry}, {d}, K);
      mfx = a; mfy = b; cx = cc; cy = d;
    }
    if (ds.heat >= 0) {   // heat_source += heat_s (model/dyn_core_nlm.F90:685-692)
      if (heat < 0) heat = ds.heat;
      else { int hn = P.val(tg + ".heat", K); P.add<S_add2>("acc_heat", {is, ie, js, je}, {heat, ds.heat}, {hn}, K); heat = hn; }
    }
    delp = ds.delp; pt = ds.pt;
    add_patch(P, "halo_delp", &mo.h_center, {delp});
    add_patch(P, "halo_pt", &mo.h_center, {pt});
    auto dzd = build_update_dz_d(P, mo, dp0, dp, c.hord_tm, rdt, zs, zh, ds.crx, ds.cry, ds.xfx, ds.yfx, tg + ".dzd", two ? c.pert.hord_tm : 0);
    ws_d = dzd.second;
    RiemOut r3 = build_riem(P, {K, 1, 0, dt, c.akap, c.ptop, c.rdgas, c.grav, c.p_fac, c.a_imp}, delp, pt, dzd.first, ds.w, dzd.second, zs, tg + ".rs3");
    const int ppe = r3.pp, zhn = r3.z;
    w = r3.w; delz = r3.dz;
    zh = zhn;
    add_patch(P, "halo_zh", &mo.h_center, {zh});
    add_patch(P, "halo_ppe", &mo.h_center, {ppe});
    // pk3 on is-2..ie+2 (solver interior + pk3_halo ring, :1129-1181), pe / peln for the remap (pe_halo :1232)
    int pk3 = P.val(tg + ".pk3", K + 1), gz3 = P.val(tg + ".gz3", K + 1);
    o.pe = P.val(tg + ".pe", K + 1); o.peln = P.val(tg + ".peln", K + 1); o.pkz = P.val(tg + ".pkz", K);
    add_col<S_geopk>(P, "pk3", {c.ptop, c.akap, c.cp_air, 2, 1, K}, {delp, pt, s.phis}, {pk3, gz3, o.pe, o.peln, o.pkz});
    o.pk = pk3;
    int gzg = P.val(tg + ".gzg", K + 1);
    P.add<S_scale>("gz_grav", {c.grav, 2}, {zh}, {gzg}, K + 1);
    int ppb = build_a2b_ord4(P, mo, ppe, K + 1, tg + ".a2b_pp"), pkb = build_a2b_ord4(P, mo, pk3, K + 1, tg + ".a2b_pk");
    int gzb = build_a2b_ord4(P, mo, gzg, K + 1, tg + ".a2b_gz"), dpb = build_a2b_ord4(P, mo, delp, K, tg + ".a2b_dp");
    u = P.val(tg + ".u", K); v = P.val(tg + ".v", K);
    if (c.beta > 0.0) {   // split_p_grad (model/dyn_core_nlm.F90:874-875); beta_d = 0 on the first sub-step (:373-375)
      int dun = P.val(tg + ".du", K), dvn = P.val(tg + ".dv", K);
      const bool first = it == 1;
      P.add<S_gradp_beta>("split_p_grad", {dt, pow(c.ptop, c.akap), first ? 0.0 : c.beta, 1, first ? 1 : 0, 0, 0},
                          {ds.u, ds.v, pkb, gzb, ppb, dpb, first ? ds.u : du_prev, first ? ds.v : dv_prev, pkb}, {u, v, dun, dvn}, K);
      du_prev = dun; dv_prev = dvn;
    } else
    P.add<S_gradp>("nh_p_grad", {dt, pow(c.ptop, c.akap), 1}, {ds.u, ds.v, pkb, gzb, ppb, dpb}, {u, v}, K);
    if (it == c.n_split) add_patch(P, "get_boundary_uv", &mo.gb_dgrid, {u, v});
    else add_patch(P, "halo_uv", &mo.h_dgrid, {u, v});
  }
  if (heat >= 0) pt = build_heat_update(P, mo, c, heat, pt, delp, delz, tag + ".heat");
  o.u = u; o.v = v; o.pt = pt; o.delp = delp; o.w = w; o.delz = delz; o.mfx = mfx; o.mfy = mfy; o.cx = cx; o.cy = cy; o.ws = ws_d;
  return o;
}

// ---------------------------------------------------------------------------------
void mod_riem(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm) {
  (void)mo;
  const int K = P.dv->g.K;
  DynConfig c; dyn_config_from(c, prm);
  const int mode = prm.geti("mode", 0);
  int delp = io.in(P, "delp", K), pt = io.in(P, "pt", K), z = io.in(P, "z", K + 1), w = io.in(P, "w", K), ws = io.in(P, "ws", 1), zb = io.in(P, "zb", 1);
  RiemOut r = build_riem(P, {K, mode, mode == 0 ? 1 : 0, prm.get("dts", 100.0), c.akap, c.ptop, c.rdgas, c.grav, c.p_fac, c.a_imp}, delp, pt, z, w, ws, zb, "riem");
  io.out(P, "pp", r.pp); io.out(P, "z_n", r.z);
  if (mode == 1) { io.out(P, "w_n", r.w); io.out(P, "dz_n", r.dz); }
}

void mod_update_dz_c(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm) {
  const int K = P.dv->g.K;
  int ut = io.in(P, "ut", K), vt = io.in(P, "vt", K), gz = io.in(P, "gz", K + 1), zs = io.in(P, "zs", 1);
  auto r = build_update_dz_c(P, mo, dp_ref_of(*prm.ak, *prm.bk, K), prm.get("dts", 100.0), zs, ut, vt, gz, "dzc");
  io.out(P, "gz_n", r.first); io.out(P, "ws", r.second);
}

void mod_update_dz_d(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm) {
  const int K = P.dv->g.K;
  DynConfig c; dyn_config_from(c, prm);
  DswParams dp; level_params(c, K, dp);
  int zh = io.in(P, "zh", K + 1), zs = io.in(P, "zs", 1), crx = io.in(P, "crx", K), cry = io.in(P, "cry", K), xfx = io.in(P, "xfx", K), yfx = io.in(P, "yfx", K);
  auto r = build_update_dz_d(P, mo, dp_ref_of(*prm.ak, *prm.bk, K), dp, c.hord_tm, 1.0 / prm.get("dts", 100.0), zs, zh, crx, cry, xfx, yfx, "dzd");
  io.out(P, "zh_n", r.first); io.out(P, "ws", r.second);
}

void mod_dyn_core_nh(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm) {
  const int K = P.dv->g.K;
  DynConfig c; dyn_config_from(c, prm); c.hydrostatic = false;
  DynState s;
  s.u = io.in(P, "u", K); s.v = io.in(P, "v", K); s.pt = io.in(P, "pt", K); s.delp = io.in(P, "delp", K);
  s.w = io.in(P, "w", K); s.delz = io.in(P, "delz", K); s.phis = io.in(P, "phis", 1);
  DynOut o = build_dyn_core_nh(P, mo, c, *prm.ak, *prm.bk, s, "dyn");
  io.out(P, "u_n", o.u); io.out(P, "v_n", o.v); io.out(P, "pt_n", o.pt); io.out(P, "delp_n", o.delp); io.out(P, "w_n", o.w); io.out(P, "delz_n", o.delz);
  io.out(P, "mfx", o.mfx); io.out(P, "mfy", o.mfy); io.out(P, "cx", o.cx); io.out(P, "cy", o.cy);
}

}  // namespace fv3lm
