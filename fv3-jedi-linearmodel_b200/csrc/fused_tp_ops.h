// Program ops of the fused fv_tp_2d kernels (tile kernels of fused_tp.h, marching kernels of fused_tp_march.h).  Included by modules.cu ONLY:
// the builders below are ordinary inline functions, so every translation unit that sees them instantiates all tile / marching kernels
// (that cost four copies of the device code and two extra minutes of build time while they lived in fused_tp.h).
#pragma once
#include "fused_tp.h"
#include "fused_tp_march.h"

namespace fv3lm {
namespace ftp {

template <template <class, bool> class K, class Fill>
void run_fused(Program& P, Op& o, int mode, bool full, const Fill& fill) {
  const Geom& g = P.dv->g;
  bool tl = false;
  if (mode == MODE_TL) for (int i : o.in) tl = tl || (P.vals[i].active && P.vals[i].pert);
  auto go = [&](auto kern) {
    kern.g = g; kern.m = P.dv->m; kern.nk = o.nk_launch;
    fill(kern, tl);
    launch_tile(kern, g.NX - 1, g.NY - 1, g.ntile * o.nk_launch);   // (the last array column / row lies outside every rectangle of fv_tp_2d)
  };
  if (tl) { if (full) go(K<Dual, true>{}); else go(K<Dual, false>{}); }
  else { if (full) go(K<double, true>{}); else go(K<double, false>{}); }
}

// appends the two fused ops to a program; the caller has already added the copy_corners patches around them.
// with_ad = false: forward sweeps only (VAR_FWD, the stage chain stays for adjoint runs); true: the ops also carry the reverse
// tile kernel (KernTpRev), the stage chain is not built at all and its six intermediates never reach HBM in any sweep.
inline void add_fused_a(Program& P, const std::string& nm, int q, int cry, int yfx, int ra_y, int crx, int fy2, int fxo,
                        const LevOrd& hord, bool full, int nk, bool with_ad) {
  Op op; op.name = nm; op.in = {q, cry, yfx, ra_y, crx}; op.out = {fy2, fxo}; op.nk_launch = nk; op.tl_only = P.tl_only;
  op.variant = with_ad ? P.variant : (int)VAR_FWD;
  for (int id : op.in) if (P.vals[id].nk != nk) throw std::runtime_error("fused fv_tp_2d: every field must have the launch's number of levels");
  op.run = [hord, full, with_ad](Program& P, Op& o, int mode) {
    if (mode == MODE_AD) {
      if (!with_ad || full) throw std::runtime_error("fused fv_tp_2d: no reverse kernel for this op");
      const Geom& g = P.dv->g;
      KernTpRev<1, false> k{};
      k.g = g; k.m = P.dv->m; k.ord = hord; k.nk = o.nk_launch;
      const Value &vq = P.vals[o.in[0]], &vci = P.vals[o.in[1]], &vfi = P.vals[o.in[2]], &vra = P.vals[o.in[3]], &vco = P.vals[o.in[4]];
      k.q = val_in(vq); k.ci = val_in(vci); k.fi = val_in(vfi); k.ra = val_in(vra); k.co = val_in(vco);
      k.aI = adj_in(P.vals[o.out[0]]); k.aO = adj_in(P.vals[o.out[1]]);
      if (!k.aI.v || !k.aO.v) throw std::runtime_error("fused fv_tp_2d: output adjoints missing");
      k.q_ad = adj_out(vq); k.ci_ad = adj_out(vci); k.fi_ad = adj_out(vfi); k.ra_ad = adj_out(vra); k.co_ad = adj_out(vco);
      launch_tile(k, g.NX - 1, g.NY - 1, g.ntile * o.nk_launch);
      return;
    }
    if (march_fwd_a(P, o, mode, full, hord)) return;
    run_fused<KernTpA>(P, o, mode, full, [&](auto& k, bool tl) {
      k.ord = hord;
      k.q = fld(P.vals[o.in[0]], tl); k.cry = fld(P.vals[o.in[1]], tl); k.yfx = fld(P.vals[o.in[2]], tl);
      k.ray = fld(P.vals[o.in[3]], tl); k.crx = fld(P.vals[o.in[4]], tl);
      k.fy2 = ofld(P.vals[o.out[0]], tl); k.fxo = ofld(P.vals[o.out[1]], tl);
    });
  };
  P.ops.push_back(op);
}
inline void add_fused_b(Program& P, const std::string& nm, int q, int crx, int xfx, int ra_x, int cry, int fy2, int fxo, int mx, int my,
                        int fx, int fy, const LevOrd& hord, bool full, int nk, bool with_ad) {
  Op op; op.name = nm; op.in = {q, crx, xfx, ra_x, cry, fy2, fxo, mx, my}; op.out = {fx, fy}; op.nk_launch = nk; op.tl_only = P.tl_only;
  op.variant = with_ad ? P.variant : (int)VAR_FWD;
  for (int id : op.in) if (P.vals[id].nk != nk) throw std::runtime_error("fused fv_tp_2d: every field must have the launch's number of levels");
  op.run = [hord, full, with_ad](Program& P, Op& o, int mode) {
    if (mode == MODE_AD) {
      if (!with_ad || full) throw std::runtime_error("fused fv_tp_2d: no reverse kernel for this op");
      const Geom& g = P.dv->g;
      KernTpRev<0, true> k{};
      k.g = g; k.m = P.dv->m; k.ord = hord; k.nk = o.nk_launch;
      const Value &vq = P.vals[o.in[0]], &vci = P.vals[o.in[1]], &vfi = P.vals[o.in[2]], &vra = P.vals[o.in[3]], &vco = P.vals[o.in[4]];
      const Value &vfy2 = P.vals[o.in[5]], &vfxo = P.vals[o.in[6]], &vmx = P.vals[o.in[7]], &vmy = P.vals[o.in[8]];
      k.q = val_in(vq); k.ci = val_in(vci); k.fi = val_in(vfi); k.ra = val_in(vra); k.co = val_in(vco);
      k.fin2 = val_in(vfy2); k.fout2 = val_in(vfxo); k.mI = val_in(vmx); k.mO = val_in(vmy);
      k.aI = adj_in(P.vals[o.out[0]]); k.aO = adj_in(P.vals[o.out[1]]);
      if (!k.aI.v || !k.aO.v) throw std::runtime_error("fused fv_tp_2d: output adjoints missing");
      k.q_ad = adj_out(vq); k.ci_ad = adj_out(vci); k.fi_ad = adj_out(vfi); k.ra_ad = adj_out(vra); k.co_ad = adj_out(vco);
      k.fin2_ad = adj_out(vfy2); k.fout2_ad = adj_out(vfxo); k.mI_ad = adj_out(vmx); k.mO_ad = adj_out(vmy);
      launch_tile(k, g.NX - 1, g.NY - 1, g.ntile * o.nk_launch);
      return;
    }
    if (march_fwd_b(P, o, mode, full, hord)) return;
    run_fused<KernTpB>(P, o, mode, full, [&](auto& k, bool tl) {
      k.ord = hord;
      k.q = fld(P.vals[o.in[0]], tl); k.crx = fld(P.vals[o.in[1]], tl); k.xfx = fld(P.vals[o.in[2]], tl); k.rax = fld(P.vals[o.in[3]], tl);
      k.cry = fld(P.vals[o.in[4]], tl); k.fy2 = fld(P.vals[o.in[5]], tl); k.fxo = fld(P.vals[o.in[6]], tl);
      k.mx = fld(P.vals[o.in[7]], tl); k.my = fld(P.vals[o.in[8]], tl);
      k.fx = ofld(P.vals[o.out[0]], tl); k.fy = ofld(P.vals[o.out[1]], tl);
    });
  };
  P.ops.push_back(op);
}

}  // namespace ftp
}  // namespace fv3lm
