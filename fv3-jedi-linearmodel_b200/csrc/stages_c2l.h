// cubed_to_latlon with c2l_ord = 4 (model/fv_grid_utils_nlm.F90:2334-2472): D-grid winds -> A-grid winds in longitude / latitude
// components, the ua / va that fv_dynamics leaves behind (fv_dynamics_nlm.F90:738) and fv3jedi_lm's step_nl returns in traj%ua, va
// (fv3jedi_lm_dynamics_mod.F90:839-840).  Trajectory-only diagnostic: not part of the increment (pert%ua = pert%va = 0, :920-921).
#pragma once
#include "engine.h"

namespace fv3lm {

// in: 0 = u, 1 = v (D grid, halo valid: mode = 1 updates it, :2366-2370), 2..5 = a11 a12 a21 a22 (2-D) ; out: 0 = ua, 1 = va.
// Interior: 4-point Lagrange interpolation (:2385-2390).  The four edge cases (:2392-2449) are one formula -- the average of the
// two winds weighted by their edge lengths -- applied to the first / last row and column of cells of a TILE.
struct S_c2l {
  static constexpr int NI = 6, NO = 2;
  struct P { int dummy; };
  static constexpr int NT = 12;
  static constexpr Tap taps[NT] = {{0, 0, -1, 0}, {0, 0, 0, 0}, {0, 0, 1, 0}, {0, 0, 2, 0}, {1, -1, 0, 0}, {1, 0, 0, 0}, {1, 1, 0, 0}, {1, 2, 0, 0},
                                   {2, 0, 0, 0}, {3, 0, 0, 0}, {4, 0, 0, 0}, {5, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P&) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie, g.js, g.je)) return;
    constexpr double c1 = 1.125, c2 = -0.125;
    T ut, vt;
    if (x.i == 1 || x.i == g.npx - 1 || x.j == 1 || x.j == g.npy - 1) {
      const double dx0 = x.M(x.m.dx), dx1 = x.M(x.m.dx, 0, 1), dy0 = x.M(x.m.dy), dy1 = x.M(x.m.dy, 1, 0);
      ut = 2.0 * (x.in(0) * dx0 + x.in(0, 0, 1) * dx1) / (dx0 + dx1);
      vt = 2.0 * (x.in(1) * dy0 + x.in(1, 1, 0) * dy1) / (dy0 + dy1);
    } else {
      ut = c2 * (x.in(0, 0, -1) + x.in(0, 0, 2)) + c1 * (x.in(0) + x.in(0, 0, 1));
      vt = c2 * (x.in(1, -1, 0) + x.in(1, 2, 0)) + c1 * (x.in(1) + x.in(1, 1, 0));
    }
    x.out(0, x.in(2) * ut + x.in(3) * vt);
    x.out(1, x.in(4) * ut + x.in(5) * vt);
  }
};

}  // namespace fv3lm
