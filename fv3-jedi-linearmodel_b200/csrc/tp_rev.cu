// reverse tile kernels of fv_tp_2d (fused_tp.h: KernTpRev) -- a translation unit of its own
#include "fused_tp.h"
#include "fused_tp_ops.h"

namespace fv3lm {
namespace ftp {

void tp_rev_a(Program& P, Op& o, const LevOrd& hord) {
      const Geom& g = P.dv->g;
      KernTpRev<1, false> k{};
      k.g = g; k.m = P.dv->m; k.ord = hord; k.nk = o.nk_launch;
      const Value &vq = P.vals[o.in[0]], &vci = P.vals[o.in[1]], &vfi = P.vals[o.in[2]], &vra = P.vals[o.in[3]], &vco = P.vals[o.in[4]];
      k.q = val_in(vq); k.ci = val_in(vci); k.fi = val_in(vfi); k.ra = val_in(vra); k.co = val_in(vco);
      k.aI = adj_in(P.vals[o.out[0]]); k.aO = adj_in(P.vals[o.out[1]]);
      if (!k.aI.v || !k.aO.v) throw std::runtime_error("fused fv_tp_2d: output adjoints missing");
      k.q_ad = adj_out(vq); k.ci_ad = adj_out(vci); k.fi_ad = adj_out(vfi); k.ra_ad = adj_out(vra); k.co_ad = adj_out(vco);
      launch_tile(k, g.NX - 1, g.NY - 1, g.ntile * o.nk_launch);
}
void tp_rev_b(Program& P, Op& o, const LevOrd& hord) {
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
}

}  // namespace ftp
}  // namespace fv3lm
