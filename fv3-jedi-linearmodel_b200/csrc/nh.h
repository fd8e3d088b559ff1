// non-hydrostatic builders (nh.cu)
#pragma once
#include "dyn.h"
#include "stages_nh.h"

namespace fv3lm {
std::pair<int, int> build_deln_public(Program& P, Mosaic& mo, int q, const LevOrd& nord, const LevD& damp, int nk, const std::string& tag);
std::pair<int, int> build_update_dz_c(Program& P, Mosaic& mo, const LevD& dp0, double dt, int zs, int ut, int vt, int gz, const std::string& tag);
std::pair<int, int> build_update_dz_d(Program& P, Mosaic& mo, const LevD& dp0, const DswParams& dp, int hord_tm, double rdt, int zs, int zh,
                                      int crx, int cry, int xfx, int yfx, const std::string& tag, int hord_tm_pert = 0);
DynOut build_dyn_core_nh(Program& P, Mosaic& mo, const DynConfig& c, const std::vector<double>& ak, const std::vector<double>& bk,
                         DynState s, const std::string& tag);
}  // namespace fv3lm
