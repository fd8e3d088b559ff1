// Program ops of the fused fv_tp_2d kernels.  The kernels are instantiated in three translation units of their own (tp_fwd.cu: forward tile
// kernels, tp_rev.cu: reverse tile kernels, tp_march.cu: row-marching forward kernels) so that the build compiles them in parallel; this
// header only declares their host entry points and holds the op builders (included by modules.cu).
#pragma once
#include "stages_tp.h"

namespace fv3lm {
namespace ftp {

// forward sweeps (NL / TL) of the two halves as tile kernels, o.in / o.out as laid out by add_fused_a / add_fused_b   (tp_fwd.cu)
void tp_fwd_a(Program& P, Op& o, int mode, bool full, const LevOrd& hord);
void tp_fwd_b(Program& P, Op& o, int mode, bool full, const LevOrd& hord);
// reverse sweeps (linear orders only)   (tp_rev.cu)
void tp_rev_a(Program& P, Op& o, const LevOrd& hord);
void tp_rev_b(Program& P, Op& o, const LevOrd& hord);
// row-marching forward kernels; false when they are switched off or do not apply (rows longer than a block)   (tp_march.cu)
bool tp_march_a(Program& P, Op& o, int mode, bool full, const LevOrd& hord);
bool tp_march_b(Program& P, Op& o, int mode, bool full, const LevOrd& hord);

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
      tp_rev_a(P, o, hord);
      return;
    }
    if (tp_march_a(P, o, mode, full, hord)) return;
    tp_fwd_a(P, o, mode, full, hord);
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
      tp_rev_b(P, o, hord);
      return;
    }
    if (tp_march_b(P, o, mode, full, hord)) return;
    tp_fwd_b(P, o, mode, full, hord);
  };
  P.ops.push_back(op);
}

}  // namespace ftp
}  // namespace fv3lm
