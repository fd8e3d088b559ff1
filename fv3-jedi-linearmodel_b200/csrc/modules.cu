#include "modules.h"
#include "stages_tp.h"
#include "fused_tp_ops.h"

namespace fv3lm {

static LevOrd lev_ord(int K, int ord, int n_sponge, int ord_sponge) {
  LevOrd o;
  for (int k = 0; k < 128; k++) o.v[k] = (signed char)enc_hord((k < n_sponge) ? ord_sponge : ord);
  (void)K;
  return o;
}

// ---------------------------------------------------------------------------------
// fv_tp_2d   (model_tlmadm/tp_core_tlm.F90:2123-2324)
// ---------------------------------------------------------------------------------
TpOut build_fv_tp_2d(Program& P, Mosaic& mo, int q, int crx, int cry, int xfx, int yfx, int ra_x, int ra_y,
                     int mfx, int mfy, const LevOrd& hord, int nk, const std::string& tag) {
  const Geom& g = P.dv->g;
  const int is = g.is, ie = g.ie, js = g.js, je = g.je, ng = g.ng;
  const int isd = is - ng, ied = ie + ng, jsd = js - ng, jed = je + ng;
  auto nm = [&](const char* s) { return tag + "." + s; };
  const char* fe = getenv("FV3LM_FUSED_TP");
  const int flevel = fe ? atoi(fe) : 2;      // default since round 2: tile kernels in every sweep (profiles/r02g_*)
  const bool fused = flevel != 0;
  // the reverse kernels cover the linear orders; a transport with a scheme of the nonlinear model has detached inputs and is never reversed
  const bool with_ad = flevel >= 2;
  int fy2 = P.val(nm("fy2"), nk), fxo = P.val(nm("fx_ou"), nk), fx = P.val(nm("fx"), nk), fy = P.val(nm("fy"), nk);
  int q_i = -1, fx2 = -1, q_j = -1, fyo = -1;
  if (!with_ad) { q_i = P.val(nm("q_i"), nk); fx2 = P.val(nm("fx2"), nk); q_j = P.val(nm("q_j"), nk); fyo = P.val(nm("fy_ou"), nk); }
  // the linear orders of the TL / AD run the lean S_ppm kernels; any order of the nonlinear model (trajectory side) S_ppm_nl
  const bool lin = ord_is_linear(hord, nk);
  auto ppm = [&](int dir, const char* nm_, const S_ppm<0>::P& p0, int qq, int cc, int out) {
    S_ppm<1>::P p1{p0.i0, p0.i1, p0.j0, p0.j1, p0.ord};
    if (dir == 0) { if (lin) P.add<S_ppm<0>>(nm_, p0, {qq, cc}, {out}, nk); else P.add<S_ppm_nl<0>>(nm_, p0, {qq, cc}, {out}, nk); }
    else { if (lin) P.add<S_ppm<1>>(nm_, p1, {qq, cc}, {out}, nk); else P.add<S_ppm_nl<1>>(nm_, p1, {qq, cc}, {out}, nk); }
  };
  // FV3LM_FUSED_TP=2 (default): every sweep runs the shared-memory-tile kernels of fused_tp.h (forward values + one reverse kernel per
  // half that recomputes the intermediates on chip); the chain is not built.  =1: forward sweeps (NL, TL) run the tile kernels, the
  // adjoint the stage chain.  =0: the eight stages in every sweep (A/B runs, parity tests of the tile kernels against the chain).
  const int var0 = P.variant;
  auto chain = [&]() { if (fused) P.variant = VAR_AD; };
  auto common = [&]() { P.variant = var0; };
  const int mx = mfx >= 0 ? mfx : xfx, my = mfy >= 0 ? mfy : yfx;
  add_patch(P, "copy_corners_y", &mo.cc2, {q});
  chain();
  if (!with_ad) {
    ppm(1, "yppm_in", {isd, ied, js, je + 1, hord}, q, cry, fy2);
    P.add<S_inner<1>>("q_i", {isd, ied, js, je}, {q, fy2, yfx, ra_y}, {q_i}, nk);
    ppm(0, "xppm_ou", {is, ie + 1, js, je, hord}, q_i, crx, fxo);
  }
  common();
  if (fused) ftp::add_fused_a(P, "tp_fused_a", q, cry, yfx, ra_y, crx, fy2, fxo, hord, !lin, nk, with_ad);
  add_patch(P, "copy_corners_x", &mo.cc1, {q});
  chain();
  if (!with_ad) {
    ppm(0, "xppm_in", {is, ie + 1, jsd, jed, hord}, q, crx, fx2);
    P.add<S_inner<0>>("q_j", {is, ie, jsd, jed}, {q, fx2, xfx, ra_x}, {q_j}, nk);
    ppm(1, "yppm_ou", {is, ie, js, je + 1, hord}, q_j, cry, fyo);
    P.add<S_favg>("fx_avg", {is, ie + 1, js, je}, {fxo, fx2, mx}, {fx}, nk);
    P.add<S_favg>("fy_avg", {is, ie, js, je + 1}, {fyo, fy2, my}, {fy}, nk);
  }
  common();
  if (fused) ftp::add_fused_b(P, "tp_fused_b", q, crx, xfx, ra_x, cry, fy2, fxo, mx, my, fx, fy, hord, !lin, nk, with_ad);
  return {fx, fy};
}

static void mod_fv_tp_2d(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm) {
  const int K = P.dv->g.K;
  int q = io.in(P, "q", K), crx = io.in(P, "crx", K), cry = io.in(P, "cry", K), xfx = io.in(P, "xfx", K),
      yfx = io.in(P, "yfx", K), ra_x = io.in(P, "ra_x", K), ra_y = io.in(P, "ra_y", K);
  int mfx = -1, mfy = -1;
  if (prm.geti("use_mf", 0)) { mfx = io.in(P, "mfx", K); mfy = io.in(P, "mfy", K); }
  LevOrd ho = lev_ord(K, prm.geti("hord", 2), prm.geti("n_sponge", 0), 1);
  TpOut o = build_fv_tp_2d(P, mo, q, crx, cry, xfx, yfx, ra_x, ra_y, mfx, mfy, ho, K, "tp");
  io.out(P, "fx", o.fx); io.out(P, "fy", o.fy);
  io.out(P, "q", q);   // corner ghost cells are rewritten in place
}

// halo exchanges and corner fills on their own (mosaic.cu)
static void mod_halo(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm) {
  const int K = P.dv->g.K;
  int q = io.in(P, "q", K), qc = io.in(P, "qc", K), u = io.in(P, "u", K), v = io.in(P, "v", K), uc = io.in(P, "uc", K), vc = io.in(P, "vc", K);
  add_patch(P, "halo_center", &mo.h_center, {q});
  add_patch(P, "halo_corner", &mo.h_corner, {qc});
  add_patch(P, "halo_dgrid", &mo.h_dgrid, {u, v});
  add_patch(P, "halo_cgrid", &mo.h_cgrid, {uc, vc});
  if (prm.geti("corners", 1)) {
    add_patch(P, "fill_corners_bgrid_x", &mo.fcb_x, {qc});
    add_patch(P, "fill_corners_dgrid", &mo.fc_dgrid_vec, {vc, uc});
    add_patch(P, "fill_4corners_x", &mo.f4c1, {q});
  }
  io.out(P, "q", q); io.out(P, "qc", qc); io.out(P, "u", u); io.out(P, "v", v); io.out(P, "uc", uc); io.out(P, "vc", vc);
}

void mod_c_sw(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm);
void mod_del2_cubed(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm);
void mod_heat_update(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm);
void mod_dyn_core(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm);
void mod_remap(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm);
void mod_step(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm);
void mod_tracer_2d(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm);
void mod_c2l_ord4(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm);
void mod_riem(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm);
void mod_update_dz_c(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm);
void mod_update_dz_d(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm);
void mod_dyn_core_nh(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm);
void mod_d_sw(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm);
void mod_a2b_ord4(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm);
void mod_geopk(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm);

struct ModEntry { const char* name; void (*fn)(Program&, Mosaic&, ModuleIO&, const ModuleParams&); const char* doc; };
static const ModEntry g_mods[] = {
    {"fv_tp_2d", mod_fv_tp_2d, "in: q crx cry xfx yfx ra_x ra_y [mfx mfy]; out: fx fy q; params: hord n_sponge use_mf"},
    {"c_sw", mod_c_sw, "in: delp pt u v w; out: delpc ptc wc uc vc ua va ut vt divg_d; params: dt2 hydrostatic nord"},
    {"d_sw", mod_d_sw, "in: delp pt u v w uc vc ua va divg_d; out: delp_n pt_n u_n v_n w_n fx fy crx cry xfx yfx; params: dt hydrostatic hord_* nord* d2_bg damp_* dddmp d4_bg (per level: name@k)"},
    {"a2b_ord4", mod_a2b_ord4, "in: qin; out: qout"},
    {"dyn_core", mod_dyn_core, "in: u v pt delp w phis; out: u_n v_n pt_n delp_n mfx mfy cx cy pkz pe peln pk; params: n_split bdt + config overrides"},
    {"remap", mod_remap, "in: pe pk peln pt q0 u v; out: pt_n q0_n u_n v_n delp_n pkz_n pe_n; params: last_step"},
    {"step", mod_step, "in: u v t delp qv ql qi o3 w phis; out: u_n v_n t_n delp_n qv_n ql_n qi_n o3_n (one fv3jedi_lm dynamics step)"},
    {"tracer_2d", mod_tracer_2d, "in: q0 q1 dp1 mfx mfy cx cy; out: q0_n q1_n; params: hord_tr q_split (0 = sub-steps from the Courant numbers) q_split_max"},
    {"c2l_ord4", mod_c2l_ord4, "in: u v a11 a12 a21 a22; out: ua va (cubed_to_latlon, c2l_ord = 4, mode = 1)"},
    {"riem", mod_riem, "in: delp pt z w ws zb; out: pp z_n [w_n dz_n]; params: mode (0 = Riem_Solver_c, 1 = Riem_Solver3) dts"},
    {"update_dz_c", mod_update_dz_c, "in: ut vt gz zs; out: gz_n ws; params: dts"},
    {"update_dz_d", mod_update_dz_d, "in: zh zs crx cry xfx yfx; out: zh_n ws; params: dts"},
    {"dyn_core_nh", mod_dyn_core_nh, "in: u v pt delp w delz phis; out: u_n v_n pt_n delp_n w_n delz_n mfx mfy cx cy"},
    {"del2_cubed", mod_del2_cubed, "in: q; out: q_n; params: cd (coefficient relative to da_min) nmax"},
    {"heat_update", mod_heat_update, "in: heat pt delp aux (pkz if hydrostatic, else delz); out: pt_n; params: hydrostatic bdt nord vtdm4 d2_bg_k1 d2_bg_k2 + constants"},
    {"halo", mod_halo, "in/out: q qc u v uc vc; params: corners"},
    {"geopk", mod_geopk, "in: delp pt phis; out: pk gz pe peln pkz; params: halo (0 = compute_fv3_pressures, 1 = C grid, 2 = D grid) cg ptop akap cp_air"},
};

void build_module(const std::string& name, Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm) {
  for (const ModEntry& e : g_mods)
    if (name == e.name) { e.fn(P, mo, io, prm); return; }
  throw std::runtime_error("unknown module " + name);
}
const char* module_list() {
  static std::string s;
  if (s.empty()) for (const ModEntry& e : g_mods) s += std::string(e.name) + ": " + e.doc + "\n";
  return s.c_str();
}

}  // namespace fv3lm
