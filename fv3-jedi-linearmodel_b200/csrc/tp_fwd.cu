// forward tile kernels of fv_tp_2d (fused_tp.h: KernTpA, KernTpB; NL and TL, linear and full sets of orders) -- a translation unit of its own
#include "fused_tp.h"
#include "fused_tp_ops.h"

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

void tp_fwd_a(Program& P, Op& o, int mode, bool full, const LevOrd& hord) {
    run_fused<KernTpA>(P, o, mode, full, [&](auto& k, bool tl) {
      k.ord = hord;
      k.q = fld(P.vals[o.in[0]], tl); k.cry = fld(P.vals[o.in[1]], tl); k.yfx = fld(P.vals[o.in[2]], tl);
      k.ray = fld(P.vals[o.in[3]], tl); k.crx = fld(P.vals[o.in[4]], tl);
      k.fy2 = ofld(P.vals[o.out[0]], tl); k.fxo = ofld(P.vals[o.out[1]], tl);
    });
}
void tp_fwd_b(Program& P, Op& o, int mode, bool full, const LevOrd& hord) {
    run_fused<KernTpB>(P, o, mode, full, [&](auto& k, bool tl) {
      k.ord = hord;
      k.q = fld(P.vals[o.in[0]], tl); k.crx = fld(P.vals[o.in[1]], tl); k.xfx = fld(P.vals[o.in[2]], tl); k.rax = fld(P.vals[o.in[3]], tl);
      k.cry = fld(P.vals[o.in[4]], tl); k.fy2 = fld(P.vals[o.in[5]], tl); k.fxo = fld(P.vals[o.in[6]], tl);
      k.mx = fld(P.vals[o.in[7]], tl); k.my = fld(P.vals[o.in[8]], tl);
      k.fx = ofld(P.vals[o.out[0]], tl); k.fy = ofld(P.vals[o.out[1]], tl);
    });
}

}  // namespace ftp
}  // namespace fv3lm
