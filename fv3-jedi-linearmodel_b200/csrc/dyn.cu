// dyn_core: the acoustic sub-cycle (model/dyn_core_nlm.F90:78-1040; TL dyn_core_tlm.F90:93,
// AD dyn_core_adm.F90:115/1686) assembled from the c_sw / d_sw / geopk / pressure-gradient
// stages with the halo exchanges of mosaic.cu in the reference's positions.
#include "dyn.h"
#include "modules.h"

namespace fv3lm {

// per-level switches of the k-loop in front of d_sw (dyn_core_nlm.F90:579-625) plus the
// TL module's first-order sponge transport (dyn_core_tlm.F90:740-926)
void level_params(const DynConfig& c, int K, DswParams& d) {
  for (int k = 0; k < 128; k++) { d.hord_mt.v[k] = d.hord_vt.v[k] = d.hord_tm.v[k] = d.hord_dp.v[k] = 2; d.nord.v[k] = d.nord_v.v[k] = d.nord_w.v[k] = d.nord_t.v[k] = 0; }
  for (int k = 0; k < 96; k++) d.d2_bg.v[k] = d.damp_v.v[k] = d.damp_w.v[k] = d.damp_t.v[k] = d.d_con.v[k] = 0.0;
  d.heat = c.d_con > 1.e-5;
  for (int k = 1; k <= K; k++) {
    int nord_k = c.nord, nord_v = std::min(2, c.nord);
    double d2 = std::min(0.20, c.d2_bg);
    double damp_vt = c.do_vort_damp ? c.vtdm4 : 0.0;
    int nord_w = nord_v, nord_t = nord_v; double damp_w = damp_vt, damp_t = damp_vt;
    double d_con_k = c.d_con;
    if (K == 1 || c.n_sponge < 0) {
      d2 = c.d2_bg;
    } else {
      if (k == 1) {
        nord_k = 0; d2 = std::max(0.01, std::max(c.d2_bg, c.d2_bg_k1)); nord_w = 0; damp_w = d2;
        if (c.do_vort_damp) { nord_v = 0; damp_vt = 0.5 * d2; }
        d_con_k = 0.0;
      } else if (k == std::max(2, c.n_sponge - 1) && c.d2_bg_k2 > 0.01) {
        nord_k = 0; d2 = std::max(c.d2_bg, c.d2_bg_k2); nord_w = 0; damp_w = d2;
        if (c.do_vort_damp) { nord_v = 0; damp_vt = 0.5 * d2; }
        d_con_k = 0.0;
      } else if (k == std::max(3, c.n_sponge) && c.d2_bg_k2 > 0.05) {
        nord_k = 0; d2 = std::max(c.d2_bg, 0.2 * c.d2_bg_k2); nord_w = 0; damp_w = d2;
        d_con_k = 0.0;
      }
    }
    // (two-sided mode: hord_ks_traj puts first-order transport into the top n_sponge_pert - 1 layers of the trajectory too,
    // dyn_core_tlm.F90:862-869)
    const bool sp = k <= c.n_sponge_ord || (c.pert.on && c.pert.hord_ks_traj && k <= c.pert.n_sponge - 1);
    d.hord_mt.v[k - 1] = sp ? 1 : enc_hord(c.hord_mt); d.hord_vt.v[k - 1] = sp ? 1 : enc_hord(c.hord_vt);
    d.hord_tm.v[k - 1] = sp ? 1 : enc_hord(c.hord_tm); d.hord_dp.v[k - 1] = sp ? 1 : enc_hord(c.hord_dp);
    d.nord.v[k - 1] = nord_k; d.nord_v.v[k - 1] = nord_v; d.nord_w.v[k - 1] = nord_w; d.nord_t.v[k - 1] = nord_t;
    d.d_con.v[k - 1] = d.heat ? d_con_k : 0.0;
    d.d2_bg.v[k - 1] = d2; d.damp_v.v[k - 1] = damp_vt; d.damp_w.v[k - 1] = damp_w; d.damp_t.v[k - 1] = damp_t;
  }
  d.dddmp = c.dddmp; d.d4_bg = c.d4_bg; d.hydrostatic = c.hydrostatic;
}

bool level_params_pert(const DynConfig& c, int K, DswParams& d) {
  if (!c.pert.on) return false;
  const DynConfig::PertSide& q = c.pert;
  level_params(c, K, d);                 // dt, hydrostatic, d_con follow the trajectory side
  d.split_damp = q.split_damp;
  d.dddmp = q.dddmp; d.d4_bg = q.d4_bg;
  for (int k = 1; k <= K; k++) {
    int hm = q.hord_mt, hv = q.hord_vt, ht = q.hord_tm, hp = q.hord_dp;
    int nord_k = q.nord, nord_v = std::min(2, q.nord);
    double d2 = std::min(0.20, q.d2_bg);
    double damp_vt = q.do_vort_damp ? q.vtdm4 : 0.0;
    int nord_w = nord_v, nord_t = nord_v; double damp_w = damp_vt, damp_t = damp_vt;
    if (k <= q.n_sponge) {               // sponge layers of the perturbation
      if (k <= q.n_sponge - 1 && q.hord_ks_pert) hm = hv = ht = hp = 1;
      nord_k = 0;
      const double dk = k == 1 ? q.d2_bg_k1 : k == 2 ? q.d2_bg_k2 : q.d2_bg_ks;
      d2 = std::max(0.01, std::max(q.d2_bg, dk));
      nord_w = 0; damp_w = d2;
      if (q.do_vort_damp) { nord_v = 0; damp_vt = 0.5 * d2; }
    }
    d.hord_mt.v[k - 1] = enc_hord(hm); d.hord_vt.v[k - 1] = enc_hord(hv); d.hord_tm.v[k - 1] = enc_hord(ht); d.hord_dp.v[k - 1] = enc_hord(hp);
    d.nord.v[k - 1] = nord_k; d.nord_v.v[k - 1] = nord_v; d.nord_w.v[k - 1] = nord_w; d.nord_t.v[k - 1] = nord_t;
    d.d2_bg.v[k - 1] = d2; d.damp_v.v[k - 1] = damp_vt; d.damp_w.v[k - 1] = damp_w; d.damp_t.v[k - 1] = damp_t;
  }
  return true;
}

int build_del2_cubed(Program& P, Mosaic& mo, int q, double cd, int nmax, int nk, const std::string& tag) {
  const int ntimes = std::min(3, nmax);
  add_patch(P, "halo_d2c", &mo.h_center, {q});
  for (int n = 1; n <= ntimes; n++) {
    const int nt = ntimes - n;
    const std::string tg = tag + ".n" + std::to_string(n);
    LevOrd lv; for (int k = 0; k < 128; k++) lv.v[k] = (signed char)nt;
    int qa = P.val(tg + ".qa", nk), fx = P.val(tg + ".fx", nk), fy = P.val(tg + ".fy", nk), qn = P.val(tg + ".q", nk);
    P.add<S_d2c_corner>("d2c_corner", {nt}, {q}, {qa}, nk);
    if (nt > 0) add_patch(P, "d2c_cc1", &mo.cc1, {qa});
    P.add<S_del_flux<0>>("d2c_fx", {lv, 0}, {qa, qa}, {fx}, nk);     // fx = del6_v (q(i-1) - q(i)) on the domain widened by nt
    if (nt > 0) add_patch(P, "d2c_cc2", &mo.cc2, {qa});
    P.add<S_del_flux<1>>("d2c_fy", {lv, 0}, {qa, qa}, {fy}, nk);
    P.add<S_d2c_upd>("d2c_upd", {nt, cd}, {qa, fx, fy}, {qn}, nk);
    q = qn;
  }
  return q;
}

int build_heat_update(Program& P, Mosaic& mo, const DynConfig& c, int heat, int pt, int delp, int aux, const std::string& tag) {
  const int K = P.dv->g.K;
  // n_con (model/dyn_core_nlm.F90:272-285; convert_ke = F)
  int n_con;
  if (c.vtdm4 > 1.e-4) n_con = K;
  else if (c.d2_bg_k1 < 1.e-3) n_con = 0;
  else n_con = c.d2_bg_k2 < 1.e-3 ? 1 : 2;
  if (n_con == 0) return pt;
  const int nf_ke = std::min(3, c.nord + 1);
  int hs = build_del2_cubed(P, mo, heat, 0.20 * P.dv->m.da_min, nf_ke, K, tag + ".d2c");
  int ptn = P.val(tag + ".pt_heat", K);
  P.add<S_heat_pt>("heat_pt", {std::min(n_con, K), c.hydrostatic ? 1 : 0, c.cp_air, c.cp_air - c.rdgas, fabs(c.bdt * c.delt_max), -c.rdgas / c.grav, c.akap / (1.0 - c.akap)},
                   {pt, hs, delp, aux}, {ptn}, K);
  return ptn;
}

DynOut build_dyn_core(Program& P, Mosaic& mo, const DynConfig& c, DynState s, const std::string& tag) {
  const Geom& g = P.dv->g;
  const int K = g.K, is = g.is, ie = g.ie, js = g.js, je = g.je, ng = g.ng;
  const int isd = is - ng, ied = ie + ng, jsd = js - ng, jed = je + ng;
  const double dt = c.bdt / c.n_split, dt2 = 0.5 * dt;
  if (!c.hydrostatic) throw std::runtime_error("build_dyn_core: non-hydrostatic path is built by build_dyn_core_nh");
  DswParams dp; level_params(c, K, dp); dp.dt = dt;
  DswParams dpp; const bool two = level_params_pert(c, K, dpp); dpp.dt = dt;
  DynOut o;
  int u = s.u, v = s.v, pt = s.pt, delp = s.delp, w = s.w;
  int mfx = -1, mfy = -1, cx = -1, cy = -1, heat = -1, du_prev = -1, dv_prev = -1;
  for (int it = 1; it <= c.n_split; it++) {
    const std::string tg = tag + ".it" + std::to_string(it);
    P.mark_segment();
    CswOut cs = build_c_sw(P, mo, delp, pt, u, v, w, dt2, true, c.nord, K, tg + ".csw");
    if (c.nord > 0) add_patch(P, "halo_divgd", &mo.h_corner, {cs.divg_d});
    int pkc = P.val(tg + ".pkc", K + 1), gz = P.val(tg + ".gzc", K + 1), pe0 = P.val(tg + ".pe_c", K + 1), pl0 = P.val(tg + ".peln_c", K + 1), pz0 = P.val(tg + ".pkz_c", K);
    add_col<S_geopk>(P, "geopk_c", {c.ptop, c.akap, c.cp_air, 1, 1, K}, {cs.delpc, cs.ptc, s.phis}, {pkc, gz, pe0, pl0, pz0});
    int uc = P.val(tg + ".uc", K), vc = P.val(tg + ".vc", K);
    P.add<S_pgrad_c>("p_grad_c", {dt2, 1}, {cs.uc, cs.vc, pkc, gz, cs.delpc}, {uc, vc}, K);
    add_patch(P, "halo_ucvc", &mo.h_cgrid, {uc, vc});
    const bool ext = c.d_ext > 0.0;
    int dpc = -1;
    if (ext) { dpc = P.val(tg + ".dpc", K); P.add<S_a2b_ord2>("a2b_ord2_delp", {0}, {delp}, {dpc}, K); }   // delp at the corners, before d_sw (:642-644)
    DswOut ds = build_d_sw(P, mo, delp, pt, u, v, w, uc, vc, cs.ua, cs.va, cs.divg_d, dp, K, tg + ".dsw", two ? &dpp : nullptr, ext);
    int divg2 = -1;
    if (ext) { divg2 = P.val(tg + ".divg2", K); add_col<S_divg2>(P, "divg2", {K, c.d_ext * P.dv->m.da_min_c}, {dpc, ds.divg}, {divg2}); }
    // flux capacitors (d_sw :913-931)
    if (mfx < 0) { mfx = ds.fx; mfy = ds.fy; cx = ds.crx; cy = ds.cry; }
    else {
      int a = P.val(tg + ".mfx", K), b = P.val(tg + ".mfy", K), cc = P.val(tg + ".cx", K), d = P.val(tg + ".cy", K);
      P.add<S_add2>("acc_mfx", {is, ie + 1, js, je}, {mfx, ds.fx}, {a}, K);
      P.add<S_add2>("acc_mfy", {is, ie, js, je + 1}, {mfy, ds.fy}, {b}, K);
      P.add<S_add2>("acc_cx", {is, ie + 1, jsd, jed}, {cx, ds.crx}, {cc}, K);
      P.add<S_add2>("acc_cy", {isd, ied, js, je + 1}, {cy, ds.cry}, {d}, K);
      mfx = a; mfy = b; cx = cc; cy = d;
    }
    if (ds.heat >= 0) {   // heat_source += heat_s (:685-692)
      if (heat < 0) heat = ds.heat;
      else { int hn = P.val(tg + ".heat", K); P.add<S_add2>("acc_heat", {is, ie, js, je}, {heat, ds.heat}, {hn}, K); heat = hn; }
    }
    delp = ds.delp; pt = ds.pt;
    add_patch(P, "halo_delp", &mo.h_center, {delp});
    add_patch(P, "halo_pt", &mo.h_center, {pt});
    int pkd = P.val(tg + ".pk", K + 1), gzd = P.val(tg + ".gz", K + 1);
    o.pe = P.val(tg + ".pe", K + 1); o.peln = P.val(tg + ".peln", K + 1); o.pkz = P.val(tg + ".pkz", K);
    add_col<S_geopk>(P, "geopk_d", {c.ptop, c.akap, c.cp_air, 2, 0, K}, {delp, pt, s.phis}, {pkd, gzd, o.pe, o.peln, o.pkz});
    o.pk = pkd;
    int pkb = build_a2b_ord4(P, mo, pkd, K + 1, tg + ".a2b_pk"), gzb = build_a2b_ord4(P, mo, gzd, K + 1, tg + ".a2b_gz");
    u = P.val(tg + ".u", K); v = P.val(tg + ".v", K);
    if (c.beta > 0.0 || ext) {   // grad1_p_update (:865-866); beta_d = 0 on the first sub-step (:373-375), du / dv go to the next one; d_ext > 0 without beta: one_grad_p with its wk1 / wk2 terms
      int dun = P.val(tg + ".du", K), dvn = P.val(tg + ".dv", K);
      const bool first = it == 1 || !(c.beta > 0.0);
      P.add<S_gradp_beta>(c.beta > 0.0 ? "grad1_p_update" : "one_grad_p_ext", {dt, pow(c.ptop, c.akap), first ? 0.0 : c.beta, 0, first ? 1 : 0, ext ? 1 : 0, c.beta > 0.0 ? 1 : 0},
                          {ds.u, ds.v, pkb, gzb, pkb, pkb, first ? ds.u : du_prev, first ? ds.v : dv_prev, ext ? divg2 : pkb}, {u, v, dun, dvn}, K);
      du_prev = dun; dv_prev = dvn;
    } else
    P.add<S_gradp>("one_grad_p", {dt, pow(c.ptop, c.akap), 0}, {ds.u, ds.v, pkb, gzb, pkb, pkb}, {u, v}, K);
    if (it == c.n_split) add_patch(P, "get_boundary_uv", &mo.gb_dgrid, {u, v});
    else add_patch(P, "halo_uv", &mo.h_dgrid, {u, v});
  }
  if (heat >= 0) pt = build_heat_update(P, mo, c, heat, pt, delp, o.pkz, tag + ".heat");
  o.u = u; o.v = v; o.pt = pt; o.delp = delp; o.w = w; o.mfx = mfx; o.mfy = mfy; o.cx = cx; o.cy = cy;
  return o;
}

void dyn_config_from(DynConfig& c, const ModuleParams& prm) {
  const fv3lm_config* f = prm.cfg;
  c.hydrostatic = prm.geti("hydrostatic", f->hydrostatic) != 0;
  c.n_split = prm.geti("n_split", f->n_split);
  c.bdt = prm.get("bdt", f->dt);
  c.nord = prm.geti("nord", f->nord);
  c.hord_mt = prm.geti("hord_mt", f->hord_mt); c.hord_vt = prm.geti("hord_vt", f->hord_vt);
  c.hord_tm = prm.geti("hord_tm", f->hord_tm); c.hord_dp = prm.geti("hord_dp", f->hord_dp);
  c.n_sponge = prm.geti("n_sponge", f->n_sponge); c.n_sponge_ord = prm.geti("n_sponge_ord", 0);
  c.d2_bg = prm.get("d2_bg", f->d2_bg); c.d2_bg_k1 = prm.get("d2_bg_k1", f->d2_bg_k1); c.d2_bg_k2 = prm.get("d2_bg_k2", f->d2_bg_k2);
  c.d4_bg = prm.get("d4_bg", f->d4_bg); c.dddmp = prm.get("dddmp", f->dddmp); c.vtdm4 = prm.get("vtdm4", f->vtdm4);
  c.do_vort_damp = prm.geti("do_vort_damp", f->do_vort_damp) != 0;
  c.ptop = prm.get("ptop", f->ptop); c.akap = prm.get("akap", f->kappa); c.cp_air = prm.get("cp_air", f->cp);
  c.rdgas = prm.get("rdgas", f->rdgas); c.grav = prm.get("grav", f->grav);
  c.zvir = prm.get("zvir", f->zvir); c.k_split = prm.geti("k_split", f->k_split); c.nq = prm.geti("nq", f->nq);
  c.hord_tr = prm.geti("hord_tr", f->hord_tr);
  // 0 in the config = "not set": the library defaults (fully implicit SIM1 solver, p_fac = 0.05)
  c.a_imp = prm.get("a_imp", f->a_imp != 0.0 ? f->a_imp : 1.0);
  c.p_fac = prm.get("p_fac", f->p_fac != 0.0 ? f->p_fac : 0.05);
  c.beta = prm.get("beta", f->beta);
  c.d_ext = prm.get("d_ext", f->d_ext);
  c.d_con = prm.get("d_con", f->d_con);
  c.q_split = prm.geti("q_split", f->q_split_dynamic ? 0 : 1);
  c.q_split_max = prm.geti("q_split_max", f->q_split_max > 0 ? f->q_split_max : 3);
  if (prm.geti("two_sided", f->two_sided)) {
    // the main fields are the perturbation model's switches; the trajectory side comes from cfg.traj (module params: "t.<name>")
    DynConfig::PertSide& q = c.pert;
    q.on = true;
    q.split_damp = prm.geti("split_damp", f->split_damp) != 0;
    q.hord_ks_pert = prm.geti("hord_ks_pert", f->hord_ks_pert) != 0; q.hord_ks_traj = prm.geti("hord_ks_traj", f->hord_ks_traj) != 0;
    q.hord_mt = c.hord_mt; q.hord_vt = c.hord_vt; q.hord_tm = c.hord_tm; q.hord_dp = c.hord_dp; q.hord_tr = c.hord_tr;
    q.nord = c.nord; q.n_sponge = c.n_sponge; q.do_vort_damp = c.do_vort_damp;
    q.d2_bg = c.d2_bg; q.d2_bg_k1 = c.d2_bg_k1; q.d2_bg_k2 = c.d2_bg_k2; q.d2_bg_ks = prm.get("d2_bg_ks", f->d2_bg_ks);
    q.d4_bg = c.d4_bg; q.dddmp = c.dddmp; q.vtdm4 = c.vtdm4;
    c.hord_mt = prm.geti("t.hord_mt", f->traj.hord_mt); c.hord_vt = prm.geti("t.hord_vt", f->traj.hord_vt);
    c.hord_tm = prm.geti("t.hord_tm", f->traj.hord_tm); c.hord_dp = prm.geti("t.hord_dp", f->traj.hord_dp);
    c.hord_tr = prm.geti("t.hord_tr", f->traj.hord_tr);
    c.nord = prm.geti("t.nord", f->traj.nord); c.do_vort_damp = prm.geti("t.do_vort_damp", f->traj.do_vort_damp) != 0;
    c.n_sponge = prm.geti("t.n_sponge", f->traj.n_sponge);
    c.dddmp = prm.get("t.dddmp", f->traj.dddmp); c.d2_bg = prm.get("t.d2_bg", f->traj.d2_bg); c.d4_bg = prm.get("t.d4_bg", f->traj.d4_bg);
    c.vtdm4 = prm.get("t.vtdm4", f->traj.vtdm4); c.d2_bg_k1 = prm.get("t.d2_bg_k1", f->traj.d2_bg_k1); c.d2_bg_k2 = prm.get("t.d2_bg_k2", f->traj.d2_bg_k2);
    c.n_sponge_ord = 0;
    auto kd = [&](const char* n, int v) { int k = prm.geti(n, v); return k == 0 ? 17 : k; };     // 0 = not set -> linear
    c.kord_mt = kd("t.kord_mt", f->traj.kord_mt); c.kord_wz = kd("t.kord_wz", f->traj.kord_wz);
    c.kord_tm = kd("t.kord_tm", f->traj.kord_tm); c.kord_tr = kd("t.kord_tr", f->traj.kord_tr);
  }
}

void mod_del2_cubed(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm) {
  const int K = P.dv->g.K;
  int q = io.in(P, "q", K);
  io.out(P, "q_n", build_del2_cubed(P, mo, q, prm.get("cd", 0.20) * P.dv->m.da_min, prm.geti("nmax", 3), K, "d2c"));
}

void mod_heat_update(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm) {
  const int K = P.dv->g.K;
  DynConfig c; dyn_config_from(c, prm);
  int heat = io.in(P, "heat", K), pt = io.in(P, "pt", K), delp = io.in(P, "delp", K), aux = io.in(P, "aux", K);
  io.out(P, "pt_n", build_heat_update(P, mo, c, heat, pt, delp, aux, "heat"));
}

void mod_dyn_core(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm) {
  const int K = P.dv->g.K;
  DynConfig c; dyn_config_from(c, prm);
  DynState s;
  s.u = io.in(P, "u", K); s.v = io.in(P, "v", K); s.pt = io.in(P, "pt", K); s.delp = io.in(P, "delp", K);
  s.w = io.in(P, "w", K); s.phis = io.in(P, "phis", 1);
  DynOut o = build_dyn_core(P, mo, c, s);
  io.out(P, "u_n", o.u); io.out(P, "v_n", o.v); io.out(P, "pt_n", o.pt); io.out(P, "delp_n", o.delp);
  io.out(P, "mfx", o.mfx); io.out(P, "mfy", o.mfy); io.out(P, "cx", o.cx); io.out(P, "cy", o.cy);
  io.out(P, "pkz", o.pkz); io.out(P, "pe", o.pe); io.out(P, "peln", o.peln); io.out(P, "pk", o.pk);
}

// geopk / compute_fv3_pressures on its own (model/dyn_core_nlm.F90:1954-2087; fv3jedi_lm_dynamics_mod.F90 compute_fv3_pressures_{tlm,bwd}):
// halo = 0 is the compute-domain form the step driver uses for the output pressures, 1 / 2 the C- / D-grid calls
void mod_geopk(Program& P, Mosaic&, ModuleIO& io, const ModuleParams& prm) {
  const int K = P.dv->g.K;
  const fv3lm_config* f = prm.cfg;
  int delp = io.in(P, "delp", K), pt = io.in(P, "pt", K), phis = io.in(P, "phis", 1);
  int pk = P.val("pk", K + 1), gz = P.val("gz", K + 1), pe = P.val("pe", K + 1), peln = P.val("peln", K + 1), pkz = P.val("pkz", K);
  add_col<S_geopk>(P, "geopk", {prm.get("ptop", f->ptop), prm.get("akap", f->kappa), prm.get("cp_air", f->cp), prm.geti("halo", 0), prm.geti("cg", 0), K},
                   {delp, pt, phis}, {pk, gz, pe, peln, pkz});
  io.out(P, "pk", pk); io.out(P, "gz", gz); io.out(P, "pe", pe); io.out(P, "peln", peln); io.out(P, "pkz", pkz);
}

}  // namespace fv3lm
