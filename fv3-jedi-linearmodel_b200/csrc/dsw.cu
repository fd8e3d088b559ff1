// d_sw program builder (model/sw_core_nlm.F90:492-1545) and the shared del-n flux builder.
#include "stages_dsw.h"
#include "modules.h"
#include "fused_chain.h"

namespace fv3lm {

static int maxord(const LevOrd& o, int nk) { int m = -1; for (int k = 0; k < nk; k++) m = std::max(m, (int)o.v[k]); return m; }

// del-n damping fluxes of q (deln_flux / del6_vt_flux).  nord(k) < 0 switches a level off.
// returns {fx2, fy2}
static std::pair<int, int> build_deln(Program& P, Mosaic& mo, int q, const LevOrd& nord, const LevD& damp, int nk, const std::string& tag) {
  auto nm = [&](const std::string& s) { return tag + "." + s; };
  const int nmax = maxord(nord, nk);
  int d2 = P.val(nm("d2_0"), nk);
  P.add<S_del_d2>("del_d2", {nord, damp}, {q}, {d2}, nk);
  if (nmax > 0) add_patch(P, "del_cc1", &mo.cc1, {d2});
  int fx2 = P.val(nm("fx2_0"), nk), fy2 = P.val(nm("fy2_0"), nk);
  P.add<S_del_flux<0>>("del_fx", {nord, 0}, {d2, d2}, {fx2}, nk);
  if (nmax > 0) add_patch(P, "del_cc2", &mo.cc2, {d2});
  P.add<S_del_flux<1>>("del_fy", {nord, 0}, {d2, d2}, {fy2}, nk);
  for (int it = 1; it <= nmax; it++) {
    int d2n = P.val(nm("d2_" + std::to_string(it)), nk);
    P.add<S_del_div>("del_div", {nord, it}, {fx2, fy2}, {d2n}, nk);
    add_patch(P, "del_cc1", &mo.cc1, {d2n});
    int fxn = P.val(nm("fx2_" + std::to_string(it)), nk), fyn = P.val(nm("fy2_" + std::to_string(it)), nk);
    P.add<S_del_flux<0>>("del_fx", {nord, it}, {d2n, fx2}, {fxn}, nk);
    add_patch(P, "del_cc2", &mo.cc2, {d2n});
    P.add<S_del_flux<1>>("del_fy", {nord, it}, {d2n, fy2}, {fyn}, nk);
    fx2 = fxn; fy2 = fyn;
  }
  return {fx2, fy2};
}

// fv_tp_2d followed by the optional deln_flux (tp_core_nlm.F90:168-208)
static TpOut build_tp_damped(Program& P, Mosaic& mo, int q, int crx, int cry, int xfx, int yfx, int ra_x, int ra_y, int mfx, int mfy,
                             int mass, const LevOrd& hord, const LevOrd& nord, const LevD& damp_c, int nk, const std::string& tag) {
  TpOut o = build_fv_tp_2d(P, mo, q, crx, cry, xfx, yfx, ra_x, ra_y, mfx, mfy, hord, nk, tag);
  const double da_min = P.dv->m.da_min;
  LevOrd n2; LevD dmp; bool any = false;
  for (int k = 0; k < 128; k++) n2.v[k] = -1;
  for (int k = 0; k < 96; k++) dmp.v[k] = 0.0;
  for (int k = 0; k < nk; k++)
    if (damp_c.v[k] > 1.e-4) { n2.v[k] = nord.v[k]; dmp.v[k] = pow(damp_c.v[k] * da_min, (double)(nord.v[k] + 1)); any = true; }
  if (!any) return o;
  const bool use_mass = mass >= 0;
  LevD one; for (int k = 0; k < 96; k++) one.v[k] = 1.0;
  auto f2 = build_deln(P, mo, q, n2, use_mass ? one : dmp, nk, tag + ".deln");
  const Geom& g = P.dv->g; (void)g;
  int fx = P.val(tag + ".fxd", nk), fy = P.val(tag + ".fyd", nk);
  int m = use_mass ? mass : q;
  P.add<S_del_add<0>>("deln_add_x", {n2, dmp, use_mass ? 1 : 0}, {o.fx, f2.first, m}, {fx}, nk);
  P.add<S_del_add<1>>("deln_add_y", {n2, dmp, use_mass ? 1 : 0}, {o.fy, f2.second, m}, {fy}, nk);
  return {fx, fy};
}

static bool same_ord(const LevOrd& a, const LevOrd& b, int nk) { for (int k = 0; k < nk; k++) if (a.v[k] != b.v[k]) return false; return true; }
static bool same_lev(const LevD& a, const LevD& b, int nk) { for (int k = 0; k < nk; k++) if (a.v[k] != b.v[k]) return false; return true; }

// prm: the switches of the nonlinear model (trajectory).  pp (optional): the perturbation-side switches of the TL/AD model
// (model_tlmadm/sw_core_tlm.F90 D_SW_TLM :1047): where the two differ, the operator is evaluated twice -- with pp for the
// perturbation (linearised about the same inputs) and with prm, on detached inputs, for the trajectory -- and spliced.
DswOut build_d_sw(Program& P, Mosaic& mo, int delp, int pt, int u, int v, int w, int uc, int vc, int ua, int va, int divg_d,
                  const DswParams& prm, int nk, const std::string& tag, const DswParams* pp_, bool want_divg) {
  auto nm = [&](const char* s) { return tag + "." + s; };
  const double da_min_c = P.dv->m.da_min_c;
  const DswParams& pp = pp_ ? *pp_ : prm;
  const bool split_damp = pp_ && pp_->split_damp;
  auto D = [&](int id) { return P.detached(id); };
  auto splice = [&](int a, int b, const std::string& name) {
    int o = P.val(name, P.vals[a].nk);
    P.add<S_splice>("splice", {0}, {a, b}, {o}, P.vals[a].nk);
    return o;
  };
  DswOut o;
  // contravariant winds, Courant numbers, flux areas
  int ut0 = P.val(nm("ut0"), nk), vt0 = P.val(nm("vt0"), nk), ut = P.val(nm("ut"), nk), vt = P.val(nm("vt"), nk);
  o.crx = P.val(nm("crx"), nk); o.xfx = P.val(nm("xfx"), nk); o.cry = P.val(nm("cry"), nk); o.yfx = P.val(nm("yfx"), nk);
  int ra_x = P.val(nm("ra_x"), nk), ra_y = P.val(nm("ra_y"), nk);
  {
    // FV3LM_FUSED_CHAIN=1 (opt-in until timed on a B200): forward sweeps run the four stages as one tile kernel (fused_chain.h),
    // ut0 / vt0 stay in shared memory; adjoint runs keep the stage-by-stage ops
    const char* fe = getenv("FV3LM_FUSED_CHAIN");
    const bool fused = fe && atoi(fe) != 0;
    const int var0 = P.variant;
    if (fused) P.variant = VAR_AD;
    P.add<S_dwind1>("dwind1", {prm.dt}, {uc, vc}, {ut0, vt0}, nk);
    P.add<S_dwind2>("dwind2", {0}, {ut0, vt0, uc, vc}, {ut, vt}, nk);
    P.add<S_dcourant>("dcourant", {prm.dt}, {ut, vt}, {o.crx, o.xfx, o.cry, o.yfx}, nk);
    P.add<S_ra>("ra", {0}, {o.xfx, o.yfx}, {ra_x, ra_y}, nk);
    P.variant = var0;
    if (fused)
      ftp::add_chain<S_dwind1, S_dwind2, S_dcourant, S_ra>(
          P, "dsw_head_fused", ftp::ppack_of<S_dwind1, S_dwind2, S_dcourant, S_ra>(S_dwind1::P{prm.dt}, S_dwind2::P{0}, S_dcourant::P{prm.dt}, S_ra::P{0}),
          {{uc, vc}, {ut0, vt0, uc, vc}, {ut, vt}, {o.xfx, o.yfx}}, {{ut0, vt0}, {ut, vt}, {o.crx, o.xfx, o.cry, o.yfx}, {ra_x, ra_y}},
          {ut, vt, o.crx, o.xfx, o.cry, o.yfx, ra_x, ra_y}, nk);
  }
  // mass
  // one transport site (:1664-1682): q is transported once when both sides use the same scheme, else twice
  auto tp_site = [&](int q, int mfx, int mfy, int mass, bool same, const LevOrd& ho_p, const LevOrd& no_p, const LevD& da_p,
                     const LevOrd& ho_t, const LevOrd& no_t, const LevD& da_t, const std::string& tg) -> TpOut {
    if (same) return build_tp_damped(P, mo, q, o.crx, o.cry, o.xfx, o.yfx, ra_x, ra_y, mfx, mfy, mass, ho_t, no_t, da_t, nk, tg);
    P.tl_only = true;
    TpOut a = build_tp_damped(P, mo, q, o.crx, o.cry, o.xfx, o.yfx, ra_x, ra_y, mfx, mfy, mass, ho_p, no_p, da_p, nk, tg + ".p");
    P.tl_only = false;
    TpOut b = build_tp_damped(P, mo, D(q), D(o.crx), D(o.cry), D(o.xfx), D(o.yfx), D(ra_x), D(ra_y), D(mfx), D(mfy), D(mass), ho_t, no_t, da_t, nk, tg + ".t");
    return {splice(a.fx, b.fx, tg + ".fx"), splice(a.fy, b.fy, tg + ".fy")};
  };
  LevD nodamp; for (int k = 0; k < 96; k++) nodamp.v[k] = 0.0;
  TpOut fdp = tp_site(delp, -1, -1, -1, same_ord(prm.hord_dp, pp.hord_dp, nk) && !split_damp, pp.hord_dp, pp.nord_v, pp.damp_v,
                      prm.hord_dp, prm.nord_v, prm.damp_v, tag + ".tp_dp");
  o.fx = fdp.fx; o.fy = fdp.fy;
  // w
  const int nh = prm.hydrostatic ? 0 : 1;
  int gxw = P.val(nm("gxw"), nk), gyw = gxw, fx2w = gxw, fy2w = gxw;
  LevD dw_on; for (int k = 0; k < 96; k++) dw_on.v[k] = 0.0;
  if (nh) {
    LevOrd nw; LevD d4;
    for (int k = 0; k < 128; k++) nw.v[k] = -1;
    for (int k = 0; k < 96; k++) d4.v[k] = 0.0;
    bool any = false;
    for (int k = 0; k < nk; k++)
      if (prm.damp_w.v[k] > 1.e-5) { nw.v[k] = prm.nord_w.v[k]; d4.v[k] = pow(prm.damp_w.v[k] * da_min_c, (double)(prm.nord_w.v[k] + 1)); dw_on.v[k] = 1.0; any = true; }
    if (any) { auto f = build_deln(P, mo, w, nw, d4, nk, tag + ".del6w"); fx2w = f.first; fy2w = f.second; }
    TpOut fw = tp_site(w, o.fx, o.fy, -1, same_ord(prm.hord_vt, pp.hord_vt, nk), pp.hord_vt, pp.nord_v, nodamp, prm.hord_vt, prm.nord_v, nodamp, tag + ".tp_w");
    gxw = fw.fx; gyw = fw.fy;
  }
  // pt
  TpOut fpt = tp_site(pt, o.fx, o.fy, delp, same_ord(prm.hord_tm, pp.hord_tm, nk) && !split_damp, pp.hord_tm, pp.nord_t, pp.damp_t,
                      prm.hord_tm, prm.nord_t, prm.damp_t, tag + ".tp_pt");
  o.delp = P.val(nm("delp"), nk); o.pt = P.val(nm("pt"), nk); o.w = P.val(nm("w"), nk);
  P.add<S_dupd>("dupd", {nh, dw_on}, {delp, pt, w, o.fx, o.fy, fpt.fx, fpt.fy, gxw, gyw, fx2w, fy2w}, {o.delp, o.pt, o.w}, nk);
  // kinetic energy
  int vb = P.val(nm("vb"), nk), ub = P.val(nm("ub"), nk), ubf = P.val(nm("ubf"), nk), vbf = P.val(nm("vbf"), nk), ke = P.val(nm("ke"), nk);
  auto tpuv = [&](int dir, const LevOrd& ho, int cc, int uu, int out) {   // lean kernels for the linear orders, S_tpuv_nl otherwise
    const bool lin = ord_is_linear(ho, nk);
    if (dir == 1) { if (lin) P.add<S_tpuv<1>>("ytp_v", {ho}, {cc, uu}, {out}, nk); else P.add<S_tpuv_nl<1>>("ytp_v", {ho}, {cc, uu}, {out}, nk); }
    else { if (lin) P.add<S_tpuv<0>>("xtp_u", {ho}, {cc, uu}, {out}, nk); else P.add<S_tpuv_nl<0>>("xtp_u", {ho}, {cc, uu}, {out}, nk); }
  };
  // FV3LM_FUSED_CHAIN=1: with one linear scheme on both sides the kinetic-energy chain dvbub -> ytp_v, xtp_u -> dke (21 array passes)
  // runs as one tile kernel in forward sweeps (7 passes; vb, ub and the two transported winds stay in shared memory)
  const char* fe_ke = getenv("FV3LM_FUSED_CHAIN");
  const bool fuse_ke = fe_ke && atoi(fe_ke) != 0 && same_ord(prm.hord_mt, pp.hord_mt, nk) && ord_is_linear(prm.hord_mt, nk);
  const int var_ke = P.variant;
  if (fuse_ke) P.variant = VAR_AD;     // the stage ops up to and including dke are the adjoint's version of the chain
  P.add<S_dvbub>("dvbub", {prm.dt}, {ut, vt, uc, vc}, {vb, ub}, nk);
  if (same_ord(prm.hord_mt, pp.hord_mt, nk)) {
    tpuv(1, prm.hord_mt, vb, v, ubf);
    tpuv(0, prm.hord_mt, ub, u, vbf);
  } else {   // :1987-1997, :2059-2068
    int ya = P.val(nm("ubf.p"), nk), yb = P.val(nm("ubf.t"), nk), xa = P.val(nm("vbf.p"), nk), xb = P.val(nm("vbf.t"), nk);
    P.tl_only = true;
    tpuv(1, pp.hord_mt, vb, v, ya);
    tpuv(0, pp.hord_mt, ub, u, xa);
    P.tl_only = false;
    tpuv(1, prm.hord_mt, D(vb), D(v), yb);
    tpuv(0, prm.hord_mt, D(ub), D(u), xb);
    P.add<S_splice>("splice", {0}, {ya, yb}, {ubf}, nk);
    P.add<S_splice>("splice", {0}, {xa, xb}, {vbf}, nk);
  }
  P.add<S_dke>("dke", {prm.dt}, {vb, ubf, ub, vbf, ut, vt, u, v}, {ke}, nk);
  P.variant = var_ke;
  if (fuse_ke)
    ftp::add_chain<S_dvbub, S_tpuv<1>, S_tpuv<0>, S_dke>(
        P, "dsw_ke_fused", ftp::ppack_of<S_dvbub, S_tpuv<1>, S_tpuv<0>, S_dke>(S_dvbub::P{prm.dt}, S_tpuv<1>::P{prm.hord_mt}, S_tpuv<0>::P{prm.hord_mt}, S_dke::P{prm.dt}),
        {{ut, vt, uc, vc}, {vb, v}, {ub, u}, {vb, ubf, ub, vbf, ut, vt, u, v}}, {{vb, ub}, {ubf}, {vbf}, {ke}}, {ke}, nk);
  // relative vorticity
  int wk = P.val(nm("wk"), nk);
  P.add<S_relvort>("relvort", {0}, {u, v}, {wk}, nk);
  // divergence damping (compute_divergence_damping, :1264-1434; split_damp: :2341-2366)
  auto div_damp = [&](const DswParams& q, int u, int v, int ua, int va, int uc, int vc, int divg_d, int wk, const std::string& tg, int* divg_out) -> int {
    auto nm = [&](const char* s) { return tg + "." + s; };
    int delpc0 = P.val(nm("delpc0"), nk), vq = P.val(nm("vq0"), nk), dd = divg_d;
    bool any0 = false; int nmax = 0;
    for (int k = 0; k < nk; k++) { if (q.nord.v[k] == 0) any0 = true; nmax = std::max(nmax, (int)q.nord.v[k]); }
    if (any0) P.add<S_ddiv0>("ddiv0", {q.nord}, {u, v, ua, va, uc, vc}, {delpc0}, nk);
    if (want_divg) { *divg_out = P.val(nm("divg"), nk); P.add<S_sel_div>("sel_div", {q.nord}, {delpc0, divg_d}, {*divg_out}, nk); }
    if (nmax > 0) {
      for (int it = 1; it <= nmax; it++) {
        const bool fill_c = (nmax - it) != 0;
        int vcw = P.val(nm("dd_vc"), nk), ucw = P.val(nm("dd_uc"), nk), ddn = P.val(nm("dd"), nk);
        if (fill_c) add_patch(P, "fill_corners_bx", &mo.fcb_x, {dd});
        P.add<S_dd_grad<0>>("dd_vc", {q.nord, it}, {dd}, {vcw}, nk);
        if (fill_c) add_patch(P, "fill_corners_by", &mo.fcb_y, {dd});
        P.add<S_dd_grad<1>>("dd_uc", {q.nord, it}, {dd}, {ucw}, nk);
        if (fill_c) add_patch(P, "fill_corners_dvec", &mo.fc_dgrid_vec, {vcw, ucw});
        P.add<S_dd_div>("dd_div", {q.nord, it}, {ucw, vcw, dd}, {ddn}, nk);
        dd = ddn;
      }
      if (q.dddmp >= 1.e-5) vq = build_a2b_ord4(P, mo, wk, nk, tg + ".a2b");
    }
    int vd = P.val(nm("vd"), nk);
    P.add<S_ddamp>("ddamp", {q.nord, q.d2_bg, q.dddmp, q.d4_bg, q.dt}, {delpc0, divg_d, vq, dd}, {vd}, nk);
    return vd;
  };
  int vd;
  if (!split_damp) vd = div_damp(prm, u, v, ua, va, uc, vc, divg_d, wk, tag, &o.divg);
  else {
    int ga = -1, gb = -1;
    P.tl_only = true;
    int a = div_damp(pp, u, v, ua, va, uc, vc, divg_d, wk, tag + ".dd_p", &ga);
    P.tl_only = false;
    int b = div_damp(prm, D(u), D(v), D(ua), D(va), D(uc), D(vc), D(divg_d), D(wk), tag + ".dd_t", &gb);
    vd = splice(a, b, tag + ".vd");
    if (want_divg) o.divg = splice(ga, gb, tag + ".divg");
  }
  // vorticity transport
  int avort = P.val(nm("avort"), nk);
  P.add<S_absvort>("absvort", {0}, {wk}, {avort}, nk);
  TpOut fv = tp_site(avort, -1, -1, -1, same_ord(prm.hord_vt, pp.hord_vt, nk), pp.hord_vt, pp.nord_v, nodamp, prm.hord_vt, prm.nord_v, nodamp, tag + ".tp_vort");
  // vorticity damping: the trajectory with (nord_v, damp_v), the perturbation with the perturbation-side pair (:2436-2451)
  int ut3 = wk, vt3 = wk;
  LevD vd_on, on_traj; for (int k = 0; k < 96; k++) vd_on.v[k] = on_traj.v[k] = 0.0;
  {
    auto del6v = [&](const DswParams& q, int wk, LevD* on, const std::string& tg) -> std::pair<int, int> {
      LevOrd nv; LevD d4; bool any = false;
      for (int k = 0; k < 128; k++) nv.v[k] = -1;
      for (int k = 0; k < 96; k++) d4.v[k] = 0.0;
      for (int k = 0; k < nk; k++)
        if (q.damp_v.v[k] > 1.e-5) { nv.v[k] = q.nord_v.v[k]; d4.v[k] = pow(q.damp_v.v[k] * da_min_c, (double)(q.nord_v.v[k] + 1)); on->v[k] = 1.0; any = true; }
      if (!any) return {wk, wk};
      return build_deln(P, mo, wk, nv, d4, nk, tg);
    };
    if (same_ord(prm.nord_v, pp.nord_v, nk) && same_lev(prm.damp_v, pp.damp_v, nk)) {
      auto f = del6v(prm, wk, &vd_on, tag + ".del6v"); ut3 = f.first; vt3 = f.second;
      on_traj = vd_on;
    } else {
      // each side adds its flux only on the levels where its own damping is on (:2506-2530): the reference's defaults have it on
      // for the perturbation (do_vort_damp_pert = T) and off for the trajectory (do_vort_damp = F)
      LevD on_p; for (int k = 0; k < 96; k++) on_p.v[k] = 0.0;
      P.tl_only = true;
      auto a = del6v(pp, wk, &on_p, tag + ".del6v_p");
      P.tl_only = false;
      auto b = del6v(prm, D(wk), &on_traj, tag + ".del6v_t");
      LevMask ma, mb;
      for (int k = 0; k < 128; k++) { ma.v[k] = (k < 96 && on_p.v[k] != 0.0) ? 1 : 0; mb.v[k] = (k < 96 && on_traj.v[k] != 0.0) ? 1 : 0; }
      for (int k = 0; k < 96; k++) vd_on.v[k] = (on_p.v[k] != 0.0 || on_traj.v[k] != 0.0) ? 1.0 : 0.0;
      ut3 = P.val(tag + ".ut3", nk); vt3 = P.val(tag + ".vt3", nk);
      P.add<S_splice_lev>("splice_lev", {ma, mb}, {a.first, b.first}, {ut3}, nk);
      P.add<S_splice_lev>("splice_lev", {ma, mb}, {a.second, b.second}, {vt3}, nk);
    }
  }
  o.u = P.val(nm("u"), nk); o.v = P.val(nm("v"), nk);
  P.add<S_duv>("duv", {vd_on}, {u, v, ke, vd, fv.fx, fv.fy, ut3, vt3}, {o.u, o.v}, nk);
  if (prm.heat) {
    // the reference reads the del6_vt_flux output whether or not the vorticity damping ran (:1496-1506): without it the arrays
    // hold unrelated winds, so that combination is refused
    for (int k = 0; k < nk; k++)
      if (prm.d_con.v[k] > 1.e-5 && on_traj.v[k] == 0.0) throw std::runtime_error("d_sw: d_con > 0 needs the vorticity damping (do_vort_damp, vtdm4 > 1e-5) on every level outside the sponge");
    int ubh = P.val(nm("ubh"), nk), fyh = P.val(nm("fyh"), nk), vbh = P.val(nm("vbh"), nk), fxh = P.val(nm("fxh"), nk);
    P.add<S_dheat_edge>("dheat_edge", {0}, {u, v, ke, vd, fv.fx, fv.fy, ut3, vt3}, {ubh, fyh, vbh, fxh}, nk);
    o.heat = P.val(nm("heat_s"), nk);
    P.add<S_dheat>("dheat", {nh, dw_on, prm.d_con}, {o.delp, ubh, fyh, vbh, fxh, w, fx2w, fy2w}, {o.heat}, nk);
  }
  return o;
}

// pre: "" for the trajectory-side switches, "p." for the perturbation side (which default to the trajectory's)
void fill_dsw_params(DswParams& d, const ModuleParams& prm, int K, const std::string& pre = "") {
  auto LO = [&](const char* base0, int dflt, bool is_hord = false) {
    auto enc = [&](int v) { return is_hord ? enc_hord(v) : v; };
    const std::string bs = pre + base0; const char* base = bs.c_str();
    if (!pre.empty()) dflt = prm.geti(base0, dflt);
    LevOrd o; int v = prm.geti(base, dflt);
    for (int k = 0; k < 128; k++) o.v[k] = (signed char)enc(v);
    for (int k = 0; k < K; k++) {
      std::string key = std::string(base) + "@" + std::to_string(k), key0 = std::string(base0) + "@" + std::to_string(k);
      if (prm.v.count(key)) o.v[k] = (signed char)enc(prm.geti(key, v));
      else if (!pre.empty() && !prm.v.count(base) && prm.v.count(key0)) o.v[k] = (signed char)enc(prm.geti(key0, v));
    }
    return o;
  };
  auto LD = [&](const char* base0, double dflt) {
    const std::string bs = pre + base0; const char* base = bs.c_str();
    if (!pre.empty()) dflt = prm.get(base0, dflt);
    LevD o; double v = prm.get(base, dflt);
    for (int k = 0; k < 96; k++) o.v[k] = v;
    for (int k = 0; k < K; k++) {
      std::string key = std::string(base) + "@" + std::to_string(k), key0 = std::string(base0) + "@" + std::to_string(k);
      if (prm.v.count(key)) o.v[k] = prm.get(key, v);
      else if (!pre.empty() && !prm.v.count(base) && prm.v.count(key0)) o.v[k] = prm.get(key0, v);
    }
    return o;
  };
  d.hord_mt = LO("hord_mt", 2, true); d.hord_vt = LO("hord_vt", 2, true); d.hord_tm = LO("hord_tm", 2, true); d.hord_dp = LO("hord_dp", 2, true);
  d.nord = LO("nord", 1); d.nord_v = LO("nord_v", 1); d.nord_w = LO("nord_w", 1); d.nord_t = LO("nord_t", 1);
  d.d2_bg = LD("d2_bg", 0.015); d.damp_v = LD("damp_v", 0.0005); d.damp_w = LD("damp_w", 0.0005); d.damp_t = LD("damp_t", 0.0005);
  d.d_con = LD("d_con", 0.0);
  d.heat = false; for (int k = 0; k < K; k++) if (d.d_con.v[k] > 1.e-5) d.heat = true;
  d.dddmp = prm.get(pre + "dddmp", prm.get("dddmp", 0.2)); d.d4_bg = prm.get(pre + "d4_bg", prm.get("d4_bg", 0.15)); d.dt = prm.get("dt", 450.0);
  d.hydrostatic = prm.geti("hydrostatic", 1) != 0;
  d.split_damp = prm.geti("split_damp", 0) != 0;
}

void mod_d_sw(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm) {
  const int K = P.dv->g.K;
  int delp = io.in(P, "delp", K), pt = io.in(P, "pt", K), u = io.in(P, "u", K), v = io.in(P, "v", K), w = io.in(P, "w", K);
  int uc = io.in(P, "uc", K), vc = io.in(P, "vc", K), ua = io.in(P, "ua", K), va = io.in(P, "va", K), divg_d = io.in(P, "divg_d", K);
  DswParams d; fill_dsw_params(d, prm, K);
  bool has_pert = prm.geti("split_damp", 0) != 0;
  for (auto& kv : prm.v) if (kv.first.rfind("p.", 0) == 0) has_pert = true;
  DswParams dp; if (has_pert) fill_dsw_params(dp, prm, K, "p.");
  DswOut o = build_d_sw(P, mo, delp, pt, u, v, w, uc, vc, ua, va, divg_d, d, K, "dsw", has_pert ? &dp : nullptr);
  io.out(P, "delp_n", o.delp); io.out(P, "pt_n", o.pt); io.out(P, "u_n", o.u); io.out(P, "v_n", o.v); io.out(P, "w_n", o.w);
  io.out(P, "fx", o.fx); io.out(P, "fy", o.fy); io.out(P, "crx", o.crx); io.out(P, "cry", o.cry); io.out(P, "xfx", o.xfx); io.out(P, "yfx", o.yfx);
  if (o.heat >= 0) io.out(P, "heat", o.heat);
}

}  // namespace fv3lm

namespace fv3lm {
std::pair<int, int> build_deln_public(Program& P, Mosaic& mo, int q, const LevOrd& nord, const LevD& damp, int nk, const std::string& tag) {
  return build_deln(P, mo, q, nord, damp, nk, tag);
}
}  // namespace fv3lm
