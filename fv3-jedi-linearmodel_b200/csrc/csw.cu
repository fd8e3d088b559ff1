// c_sw program builder (model/sw_core_nlm.F90:77-486)
#include "stages_csw.h"
#include "modules.h"
#include "fused_chain.h"

namespace fv3lm {

CswOut build_c_sw(Program& P, Mosaic& mo, int delp, int pt, int u, int v, int w, double dt2, bool hydrostatic, int nord,
                  int nk, const std::string& tag) {
  auto nm = [&](const char* s) { return tag + "." + s; };
  CswOut o;
  int utmp = P.val(nm("utmp"), nk), vtmp = P.val(nm("vtmp"), nk);
  o.ua = P.val(nm("ua"), nk); o.va = P.val(nm("va"), nk);
  int uc0 = P.val(nm("uc0"), nk), vc0 = P.val(nm("vc0"), nk);
  o.ut = P.val(nm("ut"), nk); o.vt = P.val(nm("vt"), nk);
  // d2a2c_vect
  P.add<S_d2a>("d2a", {0}, {u, v}, {utmp, vtmp}, nk);
  P.add<S_uava>("uava", {0}, {utmp, vtmp}, {o.ua, o.va}, nk);
  add_patch(P, "utmp_corners", &mo.c_utmp, {utmp, vtmp});
  add_patch(P, "ua_corners", &mo.c_ua, {o.ua, o.va});
  P.add<S_a2c<0>>("a2c_x", {dt2}, {utmp, v, o.ua}, {uc0, o.ut}, nk);
  add_patch(P, "vtmp_corners", &mo.c_vtmp, {vtmp, utmp});
  add_patch(P, "va_corners", &mo.c_va, {o.va, o.ua});
  P.add<S_a2c<1>>("a2c_y", {dt2}, {vtmp, u, o.va}, {vc0, o.vt}, nk);
  o.divg_d = P.val(nm("divg_d"), nk);
  if (nord > 0) P.add<S_divg_corner>("divergence_corner", {0}, {u, v, o.ua, o.va}, {o.divg_d}, nk);
  // first-order transport of delp, pt, w
  int fx1 = P.val(nm("fx1"), nk), fx = P.val(nm("fx"), nk), fx2 = P.val(nm("fx2"), nk);
  int fy1 = P.val(nm("fy1"), nk), fy = P.val(nm("fy"), nk), fy2 = P.val(nm("fy2"), nk);
  const int nh = hydrostatic ? 0 : 1;
  add_patch(P, "fill4c_x.delp", &mo.f4c1, {delp});
  add_patch(P, "fill4c_x.pt", &mo.f4c1, {pt});
  if (nh) add_patch(P, "fill4c_x.w", &mo.f4c1, {w});
  P.add<S_cflux<0>>("cflux_x", {nh}, {o.ut, delp, pt, w}, {fx1, fx, fx2}, nk);
  add_patch(P, "fill4c_y.delp", &mo.f4c2, {delp});
  add_patch(P, "fill4c_y.pt", &mo.f4c2, {pt});
  if (nh) add_patch(P, "fill4c_y.w", &mo.f4c2, {w});
  P.add<S_cflux<1>>("cflux_y", {nh}, {o.vt, delp, pt, w}, {fy1, fy, fy2}, nk);
  o.delpc = P.val(nm("delpc"), nk); o.ptc = P.val(nm("ptc"), nk); o.wc = P.val(nm("wc"), nk);
  P.add<S_cupd>("cupd", {nh}, {delp, pt, w, fx1, fx, fx2, fy1, fy, fy2}, {o.delpc, o.ptc, o.wc}, nk);
  // KE, absolute vorticity, C-grid wind update
  int ke = P.val(nm("ke"), nk), vort = P.val(nm("vort"), nk);
  o.uc = P.val(nm("uc"), nk); o.vc = P.val(nm("vc"), nk);
  // FV3LM_FUSED_CHAIN=1 (opt-in until timed on a B200): in forward sweeps the three stages run as one tile kernel with ke and
  // vort in shared memory (fused_chain.h); adjoint runs keep the stage-by-stage ops
  const char* fe = getenv("FV3LM_FUSED_CHAIN");
  const bool fused = fe && atoi(fe) != 0;
  const int var0 = P.variant;
  if (fused) P.variant = VAR_AD;
  P.add<S_cke>("cke", {dt2}, {o.ua, o.va, uc0, vc0, u, v}, {ke}, nk);
  P.add<S_cvort>("cvort", {0}, {uc0, vc0}, {vort}, nk);
  P.add<S_cwind>("cwind", {dt2}, {uc0, vc0, u, v, vort, ke}, {o.uc, o.vc}, nk);
  P.variant = var0;
  if (fused)
    ftp::add_chain<S_cke, S_cvort, S_cwind>(P, "csw_tail_fused", ftp::ppack_of<S_cke, S_cvort, S_cwind>(S_cke::P{dt2}, S_cvort::P{0}, S_cwind::P{dt2}),
                                            {{o.ua, o.va, uc0, vc0, u, v}, {uc0, vc0}, {uc0, vc0, u, v, vort, ke}}, {{ke}, {vort}, {o.uc, o.vc}},
                                            {o.uc, o.vc}, nk);
  return o;
}

void mod_c_sw(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams& prm) {
  const int K = P.dv->g.K;
  int delp = io.in(P, "delp", K), pt = io.in(P, "pt", K), u = io.in(P, "u", K), v = io.in(P, "v", K), w = io.in(P, "w", K);
  bool hydro = prm.geti("hydrostatic", 1) != 0;
  CswOut o = build_c_sw(P, mo, delp, pt, u, v, w, prm.get("dt2", 450.0), hydro, prm.geti("nord", 1), K, "csw");
  io.out(P, "delpc", o.delpc); io.out(P, "ptc", o.ptc); io.out(P, "wc", o.wc); io.out(P, "uc", o.uc); io.out(P, "vc", o.vc);
  io.out(P, "ua", o.ua); io.out(P, "va", o.va); io.out(P, "ut", o.ut); io.out(P, "vt", o.vt); io.out(P, "divg_d", o.divg_d);
}

}  // namespace fv3lm
