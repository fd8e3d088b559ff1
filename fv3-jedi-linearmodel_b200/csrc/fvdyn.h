// fv_dynamics / remap / tracer program builders
#pragma once
#include "dyn.h"
#include "stages_remap.h"

namespace fv3lm {

struct FvState { int u, v, w, delz, pt, delp, phis; std::vector<int> q; };
struct FvOut { int u, v, w, delz, pt, delp; std::vector<int> q; };
struct RemapOut { int u, v, pt, delp, pkz, pe, pk, peln, w, delz; std::vector<int> q; };

std::vector<int> build_tracer_2d(Program& P, Mosaic& mo, std::vector<int> q, int dp1, int mfx, int mfy, int cx, int cy, int hord_tr, const std::string& tag, int hord_tr_pert = 0,
                                 int q_split = 1, int q_split_max = 3);
RemapOut build_remap(Program& P, Mosaic& mo, const DynConfig& c, const std::vector<double>& ak, const std::vector<double>& bk,
                     int pe, int pk, int peln, int pt, std::vector<int> q, int u, int v, bool last_step, const std::string& tag,
                     int delp = -1, int w = -1, int delz = -1, int ws = -1);
FvOut build_fv_dynamics(Program& P, Mosaic& mo, const DynConfig& c, const std::vector<double>& ak, const std::vector<double>& bk, FvState s);

}  // namespace fv3lm
