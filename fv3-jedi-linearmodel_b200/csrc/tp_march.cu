// row-marching forward kernels of fv_tp_2d (fused_tp_march.h; opt-in) -- a translation unit of its own
#include "fused_tp_march.h"
#include "fused_tp_ops.h"

namespace fv3lm {
namespace ftp {

bool tp_march_a(Program& P, Op& o, int mode, bool full, const LevOrd& hord) { return march_fwd_a(P, o, mode, full, hord); }
bool tp_march_b(Program& P, Op& o, int mode, bool full, const LevOrd& hord) { return march_fwd_b(P, o, mode, full, hord); }

}  // namespace ftp
}  // namespace fv3lm
