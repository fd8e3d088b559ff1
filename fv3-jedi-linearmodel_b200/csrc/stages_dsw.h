// d_sw: D-grid shallow-water step (model/sw_core_nlm.F90:492-1545; TL sw_core_tlm.F90:1047,
// AD sw_core_adm.F90:1773/3269) with xtp_u/ytp_v (:1970/:2312), del6_vt_flux (:1547),
// deln_flux (tp_core_nlm.F90:1015) and the divergence-damping block (:1264-1434).
// Per-level switches (sponge layers, dyn_core_nlm.F90:579-625) arrive as LevOrd / LevD.
#pragma once
#include "engine.h"
#include "mosaic.h"
#include "stages_tp.h"

namespace fv3lm {

struct LevD { double v[96]; };

// ---------------------------------------------------------------------------------
// contravariant winds, pass 1: everything that depends on uc, vc only (:660-677, west/east
// ut :682-689/:699-706, south/north vt :717-725/:737-743)
// in: uc vc ; out: ut0 vt0
struct S_dwind1 {
  static constexpr int NI = 2, NO = 2;
  struct P { double dt; };
  static constexpr int NT = 8;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {0, 0, -1, 0}, {0, 1, -1, 0}, {0, 1, 0, 0},
                                   {1, 0, 0, 0}, {1, -1, 0, 0}, {1, -1, 1, 0}, {1, 0, 1, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    const int isd = g.is - g.ng, ied = g.ie + g.ng, jsd = g.js - g.ng, jed = g.je + g.ng;
    if (x.in_rect(g.is - 1, g.ie + 2, jsd, jed)) {
      T uc = x.in(0);
      if (x.i == 1 || x.i == g.npx) {
        x.out(0, val(uc) * p.dt > 0.0 ? uc / x.M(x.m.sin_sg3, -1, 0) : uc / x.M(x.m.sin_sg1));
      } else if (x.j != 0 && x.j != 1 && x.j != g.npy - 1 && x.j != g.npy) {
        x.out(0, (uc - 0.25 * x.M(x.m.cosa_u) * (x.in(1, -1, 0) + x.in(1, 0, 0) + x.in(1, -1, 1) + x.in(1, 0, 1))) * x.M(x.m.rsin_u));
      }
    }
    if (x.in_rect(isd, ied, g.js - 1, g.je + 2)) {
      T vc = x.in(1);
      if (x.j == 1 || x.j == g.npy) {
        x.out(1, val(vc) * p.dt > 0.0 ? vc / x.M(x.m.sin_sg4, 0, -1) : vc / x.M(x.m.sin_sg2));
      } else {
        x.out(1, (vc - 0.25 * x.M(x.m.cosa_v) * (x.in(0, 0, -1) + x.in(0, 1, -1) + x.in(0, 0, 0) + x.in(0, 1, 0))) * x.M(x.m.rsin_v));
      }
    }
  }
};

// contravariant winds, pass 2: cells next to the tile edges and the corner 2x2 solves
// (:690-695, :708-713, :727-732, :744-749, :762-834); everything else is copied.
// in: ut0 vt0 uc vc ; out: ut vt
struct S_dwind2 {
  static constexpr int NI = 4, NO = 2;
  struct P { int dummy; };
  static constexpr int NT = 36;
#define T9(f) {f, -1, -1, 0}, {f, 0, -1, 0}, {f, 1, -1, 0}, {f, -1, 0, 0}, {f, 0, 0, 0}, {f, 1, 0, 0}, {f, -1, 1, 0}, {f, 0, 1, 0}, {f, 1, 1, 0}
  static constexpr Tap taps[NT] = {T9(0), T9(1), T9(2), T9(3)};
#undef T9
  template <class X> DEV static void eval(X& x, const P&) {
    using T = typename X::T;
    const Geom& g = x.g;
    const int isd = g.is - g.ng, ied = g.ie + g.ng, jsd = g.js - g.ng, jed = g.je + g.ng;
    const int npx = g.npx, npy = g.npy, i = x.i, j = x.j;
    // absolute-index accessors (the point is fixed inside each branch)
    auto U = [&](int ai, int aj) { return x.in(0, ai - i, aj - j); };
    auto V = [&](int ai, int aj) { return x.in(1, ai - i, aj - j); };
    auto UC = [&](int ai, int aj) { return x.in(2, ai - i, aj - j); };
    auto VC = [&](int ai, int aj) { return x.in(3, ai - i, aj - j); };
    auto CU = [&](int ai, int aj) { return x.Mabs(x.m.cosa_u, ai, aj); };
    auto CV = [&](int ai, int aj) { return x.Mabs(x.m.cosa_v, ai, aj); };
    // ---- ut
    if (x.in_rect(g.is - 1, g.ie + 2, jsd, jed)) {
      T r = x.in(0);
      bool erow = (j == 0 || j == 1 || j == npy - 1 || j == npy);
      if (erow && i >= 3 && i <= npx - 2) {
        r = UC(i, j) - 0.25 * CU(i, j) * (V(i - 1, j) + V(i, j) + V(i - 1, j + 1) + V(i, j + 1));
      } else if (erow && (i == 2 || i == npx - 1)) {
        // corner solves: ut(2,0) ut(2,1) ut(npx-1,0) ut(npx-1,1) ut(2,npy) ut(2,npy-1) ut(npx-1,npy) ut(npx-1,npy-1)
        const int s = (i == 2) ? 1 : -1;        // direction towards the tile interior in i
        const int ie_ = (i == 2) ? 1 : npx;     // the edge column of ut
        const int iv = (i == 2) ? 1 : npx - 1;  // the vt column on the edge side
        const int ivn = (i == 2) ? 2 : npx - 2; // the vt column on the interior side
        if (j == 0 || j == npy) {
          // halo-side row: damp uses cosa_v(iv, jv)
          const int jv = (j == 0) ? 0 : npy + 1;           // vc row in the halo
          const int jin = (j == 0) ? 1 : npy;              // vt edge row
          const int jo = (j == 0) ? -1 : npy + 1;          // outer ut row
          double damp = 1.0 / (1.0 - 0.0625 * CU(i, j) * CV(iv, jv));
          r = (UC(i, j) - 0.25 * CU(i, j) * (V(iv, jin) + V(ivn, jin) + V(ivn, jv) + VC(iv, jv) -
                                            0.25 * CV(iv, jv) * (U(ie_, j) + U(ie_, jo) + U(i, jo)))) * damp;
          (void)s;
        } else {
          // interior-side row j = 1 or npy-1
          const int jv = (j == 1) ? 2 : npy - 1;           // row of the coupled vt / vc
          const int je_ = (j == 1) ? 1 : npy;              // vt edge row
          const int jo = (j == 1) ? 2 : npy - 2;           // second ut row
          double damp = 1.0 / (1.0 - 0.0625 * CU(i, j) * CV(iv, jv));
          r = (UC(i, j) - 0.25 * CU(i, j) * (V(iv, je_) + V(ivn, je_) + V(ivn, jv) + VC(iv, jv) -
                                            0.25 * CV(iv, jv) * (U(ie_, j) + U(ie_, jo) + U(i, jo)))) * damp;
        }
      }
      x.out(0, r);
    }
    // ---- vt
    if (x.in_rect(isd, ied, g.js - 1, g.je + 2)) {
      T r = x.in(1);
      bool ecol = (i == 0 || i == 1 || i == npx - 1 || i == npx);
      if (ecol && j >= 3 && j <= npy - 2) {
        r = VC(i, j) - 0.25 * CV(i, j) * (U(i, j - 1) + U(i + 1, j - 1) + U(i, j) + U(i + 1, j));
      } else if (ecol && (j == 2 || j == npy - 1)) {
        const int je_ = (j == 2) ? 1 : npy;       // edge row of vt
        const int ju = (j == 2) ? 1 : npy - 1;    // ut row on the edge side
        const int jun = (j == 2) ? 2 : npy - 2;   // ut row on the interior side
        if (i == 0 || i == npx) {
          const int iu = (i == 0) ? 0 : npx + 1;  // uc column in the halo
          const int iin = (i == 0) ? 1 : npx;     // ut edge column
          const int io = (i == 0) ? -1 : npx + 1; // outer vt column
          double damp = 1.0 / (1.0 - 0.0625 * CU(iu, ju) * CV(i, j));
          r = (VC(i, j) - 0.25 * CV(i, j) * (U(iin, ju) + U(iin, jun) + U(iu, jun) + UC(iu, ju) -
                                            0.25 * CU(iu, ju) * (V(i, je_) + V(io, je_) + V(io, j)))) * damp;
        } else {
          const int iu = (i == 1) ? 2 : npx - 1;  // column of the coupled ut / uc
          const int ie_ = (i == 1) ? 1 : npx;     // ut edge column
          const int io = (i == 1) ? 2 : npx - 2;  // second vt column
          double damp = 1.0 / (1.0 - 0.0625 * CU(iu, ju) * CV(i, j));
          r = (VC(i, j) - 0.25 * CV(i, j) * (U(ie_, ju) + U(ie_, jun) + U(iu, jun) + UC(iu, ju) -
                                            0.25 * CU(iu, ju) * (V(i, je_) + V(io, je_) + V(io, j)))) * damp;
        }
      }
      x.out(1, r);
    }
  }
};

// Courant numbers and mass-flux areas (:853-892).  in: ut vt ; out: crx xfx cry yfx
struct S_dcourant {
  static constexpr int NI = 2, NO = 4;
  struct P { double dt; };
  static constexpr int NT = 2;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    const int isd = g.is - g.ng, ied = g.ie + g.ng, jsd = g.js - g.ng, jed = g.je + g.ng;
    if (x.in_rect(g.is, g.ie + 1, jsd, jed)) {
      T xf = p.dt * x.in(0);
      if (val(xf) > 0.0) { x.out(0, xf * x.M(x.m.rdxa, -1, 0)); x.out(1, x.M(x.m.dy) * xf * x.M(x.m.sin_sg3, -1, 0)); }
      else { x.out(0, xf * x.M(x.m.rdxa)); x.out(1, x.M(x.m.dy) * xf * x.M(x.m.sin_sg1)); }
    }
    if (x.in_rect(isd, ied, g.js, g.je + 1)) {
      T yf = p.dt * x.in(1);
      if (val(yf) > 0.0) { x.out(2, yf * x.M(x.m.rdya, 0, -1)); x.out(3, x.M(x.m.dx) * yf * x.M(x.m.sin_sg4, 0, -1)); }
      else { x.out(2, yf * x.M(x.m.rdya)); x.out(3, x.M(x.m.dx) * yf * x.M(x.m.sin_sg2)); }
    }
  }
};

// ra_x, ra_y (:898-907).  in: xfx yfx ; out: ra_x ra_y
struct S_ra {
  static constexpr int NI = 2, NO = 2;
  struct P { int dummy; };
  static constexpr int NT = 4;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {0, 1, 0, 0}, {1, 0, 0, 0}, {1, 0, 1, 0}};
  template <class X> DEV static void eval(X& x, const P&) {
    const Geom& g = x.g;
    const int isd = g.is - g.ng, ied = g.ie + g.ng, jsd = g.js - g.ng, jed = g.je + g.ng;
    if (x.in_rect(g.is, g.ie, jsd, jed)) x.out(0, x.M(x.m.area) + (x.in(0) - x.in(0, 1, 0)));
    if (x.in_rect(isd, ied, g.js, g.je)) x.out(1, x.M(x.m.area) + (x.in(1) - x.in(1, 0, 1)));
  }
};

// ---------------------------------------------------------------------------------
// del-n fluxes shared by deln_flux (tp_core_nlm.F90:1015) and del6_vt_flux (sw_core :1547)
// d2 = damp(k) * q on the range widened by nord(k).  in: q ; out: d2
struct S_del_d2 {
  static constexpr int NI = 1, NO = 1;
  struct P { LevOrd nord; LevD damp; };
  static constexpr int NT = 1;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    const Geom& g = x.g;
    const int n = p.nord.v[x.kk];
    if (n < 0) return;
    if (!x.in_rect(g.is - 1 - n, g.ie + 1 + n, g.js - 1 - n, g.je + 1 + n)) return;
    x.out(0, p.damp.v[x.kk] * x.in(0));
  }
};
// flux of iteration `it` (0 = first):  DIR 0: fx2 = del6_v*(d2(i-1)-d2(i)) [sign flipped for it>0]
// levels whose nord < it pass the previous flux through.  in: d2 prev ; out: flux
template <int DIR> struct S_del_flux {
  static constexpr int NI = 2, NO = 1;
  struct P { LevOrd nord; int it; };
  static constexpr int NT = 3;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {0, DIR == 0 ? -1 : 0, DIR == 0 ? 0 : -1, 0}, {1, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    const int n = p.nord.v[x.kk];
    if (n < 0) return;
    if (n < p.it) { x.out(0, x.in(1)); return; }
    const int nt = n - p.it;
    if (DIR == 0) { if (!x.in_rect(g.is - nt, g.ie + nt + 1, g.js - nt, g.je + nt)) return; }
    else { if (!x.in_rect(g.is - nt, g.ie + nt, g.js - nt, g.je + nt + 1)) return; }
    T lo = DIR == 0 ? x.in(0, -1, 0) : x.in(0, 0, -1), hi = x.in(0);
    double c = DIR == 0 ? x.M(x.m.del6_v) : x.M(x.m.del6_u);
    x.out(0, p.it == 0 ? c * (lo - hi) : c * (hi - lo));
  }
};
// d2 = div(fx2, fy2) * rarea for iteration it >= 1.  in: fx2 fy2 ; out: d2
struct S_del_div {
  static constexpr int NI = 2, NO = 1;
  struct P { LevOrd nord; int it; };
  static constexpr int NT = 4;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {0, 1, 0, 0}, {1, 0, 0, 0}, {1, 0, 1, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    const Geom& g = x.g;
    const int n = p.nord.v[x.kk];
    if (n < p.it) return;
    const int nt = n - p.it;
    if (!x.in_rect(g.is - nt - 1, g.ie + nt + 1, g.js - nt - 1, g.je + nt + 1)) return;
    x.out(0, ((x.in(0) - x.in(0, 1, 0)) + (x.in(1) - x.in(1, 0, 1))) * x.M(x.m.rarea));
  }
};
// add the damping flux to the transport flux (tp_core_nlm.F90:1138-1161)
// in: f f2 mass ; out: f_new.   mass form: f + 0.5*damp*(mass(-1)+mass)*f2
template <int DIR> struct S_del_add {
  static constexpr int NI = 3, NO = 1;
  struct P { LevOrd nord; LevD damp; int use_mass; };
  static constexpr int NT = 4;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, 0, 0, 0}, {2, DIR == 0 ? -1 : 0, DIR == 0 ? 0 : -1, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    const Geom& g = x.g;
    if (DIR == 0) { if (!x.in_rect(g.is, g.ie + 1, g.js, g.je)) return; }
    else { if (!x.in_rect(g.is, g.ie, g.js, g.je + 1)) return; }
    if (p.nord.v[x.kk] < 0) { x.out(0, x.in(0)); return; }
    if (p.use_mass) {
      auto m = (DIR == 0 ? x.in(2, -1, 0) : x.in(2, 0, -1)) + x.in(2);
      x.out(0, x.in(0) + (0.5 * p.damp.v[x.kk]) * m * x.in(1));
    } else {
      x.out(0, x.in(0) + x.in(1));
    }
  }
};

// ---------------------------------------------------------------------------------
// delp / pt / w update (:956-960, :1018-1031, :1236-1247)
// in: delp pt w fx fy gxp gyp gxw gyw fx2w fy2w ; out: delp_new pt_new w_new
struct S_dupd {
  static constexpr int NI = 11, NO = 3;
  struct P { int nonhydro; LevD dw_on; };
  static constexpr int NT = 19;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, 0, 0, 0},
                                   {3, 0, 0, 0}, {3, 1, 0, 0}, {4, 0, 0, 0}, {4, 0, 1, 0},
                                   {5, 0, 0, 0}, {5, 1, 0, 0}, {6, 0, 0, 0}, {6, 0, 1, 0},
                                   {7, 0, 0, 0}, {7, 1, 0, 0}, {8, 0, 0, 0}, {8, 0, 1, 0},
                                   {9, 0, 0, 0}, {9, 1, 0, 0}, {10, 0, 0, 0}, {10, 0, 1, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie, g.js, g.je)) return;
    double ra = x.M(x.m.rarea);
    T dp = x.in(0);
    T ptd = x.in(1) * dp + ((x.in(5) - x.in(5, 1, 0)) + (x.in(6) - x.in(6, 0, 1))) * ra;
    T dpn = dp + ((x.in(3) - x.in(3, 1, 0)) + (x.in(4) - x.in(4, 0, 1))) * ra;
    x.out(0, dpn);
    x.out(1, ptd / dpn);
    if (p.nonhydro) {
      T wd = dp * x.in(2) + ((x.in(7) - x.in(7, 1, 0)) + (x.in(8) - x.in(8, 0, 1))) * ra;
      // (levels without w damping: the del6 flux arrays were not written there -- or do not exist at all -- and are not read)
      if (p.dw_on.v[x.kk] != 0.0) x.out(2, wd / dpn + ((x.in(9) - x.in(9, 1, 0)) + (x.in(10) - x.in(10, 0, 1))) * ra);
      else x.out(2, wd / dpn);
    }
  }
};

// B-grid transporting winds for the KE (:1063-1163).  in: ut vt uc vc ; out: vb ub
struct S_dvbub {
  static constexpr int NI = 4, NO = 2;
  struct P { double dt; };
  static constexpr int NT = 12;
  static constexpr Tap taps[NT] = {{0, 0, -2, 0}, {0, 0, -1, 0}, {0, 0, 0, 0}, {0, 0, 1, 0},
                                   {1, -2, 0, 0}, {1, -1, 0, 0}, {1, 0, 0, 0}, {1, 1, 0, 0},
                                   {2, 0, -1, 0}, {2, 0, 0, 0}, {3, -1, 0, 0}, {3, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie + 1, g.js, g.je + 1)) return;
    const double dt4 = 0.25 * p.dt, dt5 = 0.5 * p.dt;
    const int i = x.i, j = x.j, npx = g.npx, npy = g.npy;
    T vb, ub;
    if (j == 1 || j == npy) vb = dt5 * (x.in(1, -1, 0) + x.in(1));
    else if (i == 1 || i == npx) vb = dt4 * (-x.in(1, -2, 0) + 3.0 * (x.in(1, -1, 0) + x.in(1)) - x.in(1, 1, 0));
    else vb = dt5 * ((x.in(3, -1, 0) + x.in(3)) - (x.in(2, 0, -1) + x.in(2)) * x.M(x.m.cosa)) * x.M(x.m.rsina);
    if (i == 1 || i == npx) ub = dt5 * (x.in(0, 0, -1) + x.in(0));
    else if (j == 1 || j == npy) ub = dt4 * (-x.in(0, 0, -2) + 3.0 * (x.in(0, 0, -1) + x.in(0)) - x.in(0, 0, 1));
    else ub = dt5 * ((x.in(2, 0, -1) + x.in(2)) - (x.in(3, -1, 0) + x.in(3)) * x.M(x.m.cosa)) * x.M(x.m.rsina);
    x.out(0, vb);
    x.out(1, ub);
  }
};

// xtp_u / ytp_v (linear orders).  DIR 0 = xtp_u: in: c(=ub) u ; out: flux.   (:1970-2100)
// FULL = false: linear orders only (1, 2, 333; the TL / AD kernels); FULL = true (S_tpuv_nl): every order of the nonlinear model
template <int DIR, bool FULL = false> struct S_tpuv {
  static constexpr int NI = 2, NO = 1;
  struct P { LevOrd ord; };
  static constexpr int NT = 7;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0},
                                   {1, DIR == 0 ? -3 : 0, DIR == 0 ? 0 : -3, 0}, {1, DIR == 0 ? -2 : 0, DIR == 0 ? 0 : -2, 0},
                                   {1, DIR == 0 ? -1 : 0, DIR == 0 ? 0 : -1, 0}, {1, 0, 0, 0},
                                   {1, DIR == 0 ? 1 : 0, DIR == 0 ? 0 : 1, 0},   {1, DIR == 0 ? 2 : 0, DIR == 0 ? 0 : 2, 0}};
  template <class X> DEV static typename X::T Q(const X& x, int d) { return DIR == 0 ? x.in(1, d, 0) : x.in(1, 0, d); }
  template <class X> DEV static double DD(const X& x, int d) { return DIR == 0 ? x.M(x.m.dx, d, 0) : x.M(x.m.dy, 0, d); }
  // face value between cells (pos+d-1) and (pos+d)
  template <class X> DEV static typename X::T face(const X& x, int d) {
    using T = typename X::T;
    const int ia = (DIR == 0 ? x.i : x.j) + d;
    const int np = DIR == 0 ? x.g.npx : x.g.npy;
    if (ia == 0 || ia == np - 1) return tp::c1 * Q(x, d - 2) + tp::c2 * Q(x, d - 1) + tp::c3 * Q(x, d);
    if (ia == 2 || ia == np + 1) return tp::c3 * Q(x, d - 1) + tp::c2 * Q(x, d) + tp::c1 * Q(x, d + 1);
    if (ia == 1 || ia == np) {
      double a0 = DD(x, d - 1), am = DD(x, d - 2), a1 = DD(x, d), a2 = DD(x, d + 1);
      T l = ((2.0 * a0 + am) * Q(x, d - 1) - a0 * Q(x, d - 2)) / (a0 + am);
      T r = ((2.0 * a1 + a2) * Q(x, d) - a1 * Q(x, d + 1)) / (a1 + a2);
      return 0.5 * (l + r);
    }
    return tp::p1 * (Q(x, d - 1) + Q(x, d)) + tp::p2 * (Q(x, d - 2) + Q(x, d + 1));
  }
  // bl, br of the cell at offset co (tile index ic) for iord >= 8; edge_row: the row / column is a cube edge (bl = br = 0 next to it)
  template <class X> DEV static void mono(const X& x, int co, int ic, int np, bool edge_row, int ord, typename X::T& bl, typename X::T& br) {
    using T = typename X::T;
    using namespace tp;
    constexpr double nz = 1.e-9;     // near_zero of sw_core (:37)
    auto q = [&](int d) { return Q(x, co + d); };
    auto dq = [&](int d) { return q(d + 1) - q(d); };
    auto dm = [&](int d) {
      T qm = q(d - 1), q0 = q(d), qp = q(d + 1);
      T xt = 0.25 * (qp - qm);
      return sgn_of(min3(m_abs(xt), max3(qm, q0, qp) - q0, q0 - min3(qm, q0, qp)), xt);
    };
    auto al = [&](int d) { return 0.5 * (q(d - 1) + q(d)) + r3 * (dm(d - 1) - dm(d)); };
    auto two = [&](int d) {          // x0L + x0R at the cube edge between cells ic + d - 1 and ic + d
      const double am = DD(x, co + d - 2), a0 = DD(x, co + d - 1), a1 = DD(x, co + d), a2 = DD(x, co + d + 1);
      return 0.5 * ((2.0 * a0 + am) * q(d - 1) - a0 * q(d - 2)) / (a0 + am) + 0.5 * ((2.0 * a1 + a2) * q(d) - a1 * q(d + 1)) / (a1 + a2);
    };
    const T q0 = q(0);
    if (ic >= 3 && ic <= np - 3) {
      if (ord == 8) {
        T xt = 2.0 * dm(0);
        bl = T(0.0) - sgn_of(m_min(m_abs(xt), m_abs(al(0) - q0)), xt);
        br = sgn_of(m_min(m_abs(xt), m_abs(al(1) - q0)), xt);
      } else if (ord == 9 || ord == 10) {
        bl = al(0) - q0; br = al(1) - q0;
        bool lim = true;
        if (ord == 10) {
          lim = false;
          if (val(m_abs(dm(0))) < nz) { if (val(m_abs(dm(-1))) + val(m_abs(dm(1))) < nz) { bl = T(0.0); br = T(0.0); } }
          else if (fabs(3.0 * (val(bl) + val(br))) > fabs(val(bl) - val(br))) lim = true;
        }
        if (lim) {
          T pmp_1 = T(0.0) - 2.0 * dq(0), lac_1 = pmp_1 + 1.5 * dq(1);
          bl = m_min(max3(T(0.0), pmp_1, lac_1), m_max(bl, min3(T(0.0), pmp_1, lac_1)));
          T pmp_2 = 2.0 * dq(-1), lac_2 = pmp_2 - 1.5 * dq(-2);
          br = m_min(max3(T(0.0), pmp_2, lac_2), m_max(br, min3(T(0.0), pmp_2, lac_2)));
        }
      } else { bl = al(0) - q0; br = al(1) - q0; }
      return;
    }
    if (ic == 2) { bl = (s15 * q(-1) + s11 * q0 - s14 * dm(0)) - q0; br = al(1) - q0; pert_ppm(q0, bl, br, -1); }
    else if (ic == np - 2) { bl = al(0) - q0; br = (s15 * q(1) + s11 * q0 + s14 * dm(0)) - q0; pert_ppm(q0, bl, br, -1); }
    else if (edge_row) { bl = T(0.0); br = T(0.0); }
    else if (ic == 1) { bl = two(0) - q0; br = (s15 * q0 + s11 * q(1) - s14 * dm(1)) - q0; }
    else if (ic == 0) { bl = s14 * dm(-1) - s11 * dq(-1); br = two(1) - q0; }
    else if (ic == np - 1) { bl = (s15 * q0 + s11 * q(-1) + s14 * dm(-1)) - q0; br = two(1) - q0; }
    else { bl = two(0) - q0; br = s11 * dq(0) - s14 * dm(1); }    // ic == np
  }
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie + 1, g.js, g.je + 1)) return;
    T c = x.in(0);
    const int ord = p.ord.v[x.kk];
    const bool pos = val(c) > 0.0;
    const int dc = pos ? -1 : 0;                 // upwind cell offset
    T uc = Q(x, dc);
    if (ord == 1) { x.out(0, uc); return; }
    if (ord == ORD333) {   // sw_core_tlm.F90:7332-7356 (xtp_u), :7552-7577 (ytp_v): no cube-edge special cases
      double rd3 = DIR == 0 ? x.M(x.m.rdx, dc, 0) : x.M(x.m.rdy, 0, dc);
      T cf = c * rd3, um1 = Q(x, -1), u0 = Q(x, 0);
      if (pos) { T um2 = Q(x, -2); x.out(0, (2.0 * u0 + 5.0 * um1 - um2) / 6.0 - 0.5 * cf * (u0 - um1) + cf * cf / 6.0 * (u0 - 2.0 * um1 + um2)); }
      else { T u1 = Q(x, 1); x.out(0, (2.0 * um1 + 5.0 * u0 - u1) / 6.0 - 0.5 * cf * (u0 - um1) + cf * cf / 6.0 * (u1 - 2.0 * u0 + um1)); }
      return;
    }
    const int ic = (DIR == 0 ? x.i : x.j) + dc;  // upwind cell index along the sweep
    const int np = DIR == 0 ? g.npx : g.npy;
    const int jt = DIR == 0 ? x.j : x.i;         // index across the sweep
    const int npt = DIR == 0 ? g.npy : g.npx;
    T bl, br;
    if constexpr (FULL) {
    if (ord >= 8 && ord <= 13) {
      // monotone schemes of the nonlinear model (sw_core_nlm.F90 xtp_u :2166-2306, ytp_v :2548-2735): iord 8, 9, 10, else unlimited
      mono(x, dc, ic, np, jt == 1 || jt == npt, ord, bl, br);
      double rdm = DIR == 0 ? x.M(x.m.rdx, dc, 0) : x.M(x.m.rdy, 0, dc);
      T cf = c * rdm;
      x.out(0, pos ? uc + (1.0 - cf) * (br - cf * (bl + br)) : uc + (1.0 + cf) * (bl + cf * (bl + br)));
      return;
    }
    if (ord >= 3 && ord <= 7) {
      // smoothness-switch schemes of the nonlinear model (sw_core_nlm.F90 xtp_u :2082-2160, ytp_v): need bl, br of both cells
      auto lin = [&](int co, T& l, T& r) {
        const int icc = (DIR == 0 ? x.i : x.j) + co;
        T uu = Q(x, co);
        if ((icc == 0 || icc == 1 || icc == np - 1 || icc == np) && (jt == 1 || jt == npt)) { l = T(0.0); r = T(0.0); }
        else { l = face(x, co) - uu; r = face(x, co + 1) - uu; }
      };
      T blm, brm, blp, brp; lin(-1, blm, brm); lin(0, blp, brp);
      T b0m = blm + brm, b0p = blp + brp, um = Q(x, -1), up = Q(x, 0);
      const double rdc = DIR == 0 ? x.M(x.m.rdx, dc, 0) : x.M(x.m.rdy, 0, dc);
      T cf = c * rdc;
      bool s5m, s5p, s6m = false, s6p = false;
      if (ord <= 4) {
        s5m = fabs(val(b0m)) < fabs(val(blm) - val(brm)); s6m = 3.0 * fabs(val(b0m)) < fabs(val(blm) - val(brm));
        s5p = fabs(val(b0p)) < fabs(val(blp) - val(brp)); s6p = 3.0 * fabs(val(b0p)) < fabs(val(blp) - val(brp));
        if (ord == 3) {
          T f0 = T(0.0);
          if (pos) { if (s6m || s5p) f0 = brm - cf * b0m; else if (s5m) f0 = tp::sgn_of(m_min(m_abs(blm), m_abs(brm)), brm); x.out(0, um + (1.0 - cf) * f0); }
          else { if (s6p || s5m) f0 = blp + cf * b0p; else if (s5p) f0 = tp::sgn_of(m_min(m_abs(blp), m_abs(brp)), blp); x.out(0, up + (1.0 + cf) * f0); }
          return;
        }
        if (pos) x.out(0, (s6m || s5p) ? um + (1.0 - cf) * (brm - cf * b0m) : um);
        else x.out(0, (s6p || s5m) ? up + (1.0 + cf) * (blp + cf * b0p) : up);
        return;
      }
      if (ord == 5) { s5m = val(blm) * val(brm) < 0.0; s5p = val(blp) * val(brp) < 0.0; }
      else { s5m = fabs(3.0 * val(b0m)) < fabs(val(blm) - val(brm)); s5p = fabs(3.0 * val(b0p)) < fabs(val(blp) - val(brp)); }
      (void)s6m; (void)s6p;
      if (pos) x.out(0, (s5m || s5p) ? um + (1.0 - cf) * (brm - cf * b0m) : um);
      else x.out(0, (s5m || s5p) ? up + (1.0 + cf) * (blp + cf * b0p) : up);
      return;
    }
    }   // FULL
    if ((ic == 0 || ic == 1 || ic == np - 1 || ic == np) && (jt == 1 || jt == npt)) { bl = T(0.0); br = T(0.0); }
    else { bl = face(x, dc) - uc; br = face(x, dc + 1) - uc; }
    T b0 = bl + br;
    double rd = DIR == 0 ? x.M(x.m.rdx, dc, 0) : x.M(x.m.rdy, 0, dc);
    T cfl = c * rd;
    x.out(0, pos ? uc + (1.0 - cfl) * (br - cfl * b0) : uc + (1.0 + cfl) * (bl + cfl * b0));
  }
};

template <int DIR> using S_tpuv_nl = S_tpuv<DIR, true>;

// kinetic energy at corners (:1111-1202).  in: vb ubf ub vbf ut vt u v ; out: ke
struct S_dke {
  static constexpr int NI = 8, NO = 1;
  struct P { double dt; };
  static constexpr int NT = 12;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, 0, 0, 0}, {3, 0, 0, 0},
                                   {4, 0, 0, 0}, {4, 0, -1, 0}, {5, 0, 0, 0}, {5, -1, 0, 0},
                                   {6, 0, 0, 0}, {6, -1, 0, 0}, {7, 0, 0, 0}, {7, 0, -1, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie + 1, g.js, g.je + 1)) return;
    const int i = x.i, j = x.j, npx = g.npx, npy = g.npy;
    const double dt6 = p.dt / 6.0;
    T ke;
    if (i == 1 && j == 1)
      ke = dt6 * ((x.in(4) + x.in(4, 0, -1)) * x.in(6) + (x.in(5) + x.in(5, -1, 0)) * x.in(7) + (x.in(4) + x.in(5)) * x.in(6, -1, 0));
    else if (i == npx && j == 1)
      ke = dt6 * ((x.in(4) + x.in(4, 0, -1)) * x.in(6, -1, 0) + (x.in(5) + x.in(5, -1, 0)) * x.in(7) + (x.in(4) - x.in(5, -1, 0)) * x.in(6));
    else if (i == npx && j == npy)
      ke = dt6 * ((x.in(4) + x.in(4, 0, -1)) * x.in(6, -1, 0) + (x.in(5) + x.in(5, -1, 0)) * x.in(7, 0, -1) + (x.in(4, 0, -1) + x.in(5, -1, 0)) * x.in(6));
    else if (i == 1 && j == npy)
      ke = dt6 * ((x.in(4) + x.in(4, 0, -1)) * x.in(6) + (x.in(5) + x.in(5, -1, 0)) * x.in(7, 0, -1) + (x.in(4, 0, -1) - x.in(5)) * x.in(6, -1, 0));
    else
      ke = 0.5 * (x.in(0) * x.in(1) + x.in(2) * x.in(3));
    x.out(0, ke);
  }
};

// relative vorticity on the A grid (:1204-1221).  in: u v ; out: wk
struct S_relvort {
  static constexpr int NI = 2, NO = 1;
  struct P { int dummy; };
  static constexpr int NT = 4;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {0, 0, 1, 0}, {1, 0, 0, 0}, {1, 1, 0, 0}};
  template <class X> DEV static void eval(X& x, const P&) {
    const Geom& g = x.g;
    if (!x.in_rect(g.is - g.ng, g.ie + g.ng, g.js - g.ng, g.je + g.ng)) return;
    x.out(0, x.M(x.m.rarea) * ((x.in(0) * x.M(x.m.dx) - x.in(0, 0, 1) * x.M(x.m.dx, 0, 1)) +
                               (x.in(1, 1, 0) * x.M(x.m.dy, 1, 0) - x.in(1) * x.M(x.m.dy))));
  }
};

// nord == 0 divergence (:1284-1340), levels with nord(k) == 0 only.  in: u v ua va uc vc ; out: delpc
struct S_ddiv0 {
  static constexpr int NI = 6, NO = 1;
  struct P { LevOrd nord; };
  static constexpr int NT = 12;   // uc, vc only select the upwind sin_sg: no derivative, no taps
  static constexpr Tap taps[NT] = {{0, -1, 0, 0}, {0, 0, 0, 0}, {1, 0, -1, 0}, {1, 0, 0, 0},
                                   {2, -1, -1, 0}, {2, 0, -1, 0}, {2, -1, 0, 0}, {2, 0, 0, 0},
                                   {3, -1, -1, 0}, {3, -1, 0, 0}, {3, 0, -1, 0}, {3, 0, 0, 0}};
  template <class X> DEV static typename X::T ptc(const X& x, int di) {   // ptc(i+di, j)
    if (x.j == 1 || x.j == x.g.npy)
      return x.in(0, di, 0) * x.M(x.m.dyc, di, 0) * (val(x.in(5, di, 0)) > 0.0 ? x.M(x.m.sin_sg4, di, -1) : x.M(x.m.sin_sg2, di, 0));
    return (x.in(0, di, 0) - 0.5 * (x.in(3, di, -1) + x.in(3, di, 0)) * x.M(x.m.cosa_v, di, 0)) * x.M(x.m.dyc, di, 0) * x.M(x.m.sina_v, di, 0);
  }
  template <class X> DEV static typename X::T vo(const X& x, int dj) {    // vort(i, j+dj)
    if (x.i == 1 || x.i == x.g.npx)
      return x.in(1, 0, dj) * x.M(x.m.dxc, 0, dj) * (val(x.in(4, 0, dj)) > 0.0 ? x.M(x.m.sin_sg3, -1, dj) : x.M(x.m.sin_sg1, 0, dj));
    return (x.in(1, 0, dj) - 0.5 * (x.in(2, -1, dj) + x.in(2, 0, dj)) * x.M(x.m.cosa_u, 0, dj)) * x.M(x.m.dxc, 0, dj) * x.M(x.m.sina_u, 0, dj);
  }
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (p.nord.v[x.kk] != 0) return;
    if (!x.in_rect(g.is, g.ie + 1, g.js, g.je + 1)) return;
    T vm = vo(x, -1), v0 = vo(x, 0);
    T d = (vm - v0) + (ptc(x, -1) - ptc(x, 0));
    if ((x.i == 1 || x.i == g.npx) && x.j == 1) d = d - vm;
    if ((x.i == 1 || x.i == g.npx) && x.j == g.npy) d = d + v0;
    x.out(0, x.M(x.m.rarea_c) * d);
  }
};

// the divergence d_sw hands back in delpc (the caller's vt, model/dyn_core_nlm.F90:656): its own nord = 0 divergence (:1340) on the
// levels with nord(k) = 0, a copy of the C-grid step's divg_d (:1353) elsewhere.  in: delpc0 divg_d ; out: divg
struct S_sel_div {
  static constexpr int NI = 2, NO = 1;
  struct P { LevOrd nord; };
  static constexpr int NT = 2;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie + 1, g.js, g.je + 1)) return;
    if (p.nord.v[x.kk] == 0) x.out(0, x.in(0)); else x.out(0, x.in(1));
  }
};

// higher-order divergence damping iteration (:1358-1400)
// vc = (divg(i+1,j) - divg(i,j)) * divg_u ; uc = (divg(i,j+1) - divg(i,j)) * divg_v.  in: dd ; out: w
template <int DIR> struct S_dd_grad {
  static constexpr int NI = 1, NO = 1;
  struct P { LevOrd nord; int it; };
  static constexpr int NT = 2;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {0, DIR == 0 ? 1 : 0, DIR == 0 ? 0 : 1, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    const Geom& g = x.g;
    const int n = p.nord.v[x.kk];
    if (n < p.it) return;
    const int nt = n - p.it;
    if (DIR == 0) { if (!x.in_rect(g.is - 1 - nt, g.ie + 1 + nt, g.js - nt, g.je + 1 + nt)) return; }
    else { if (!x.in_rect(g.is - nt, g.ie + 1 + nt, g.js - 1 - nt, g.je + 1 + nt)) return; }
    if (DIR == 0) x.out(0, (x.in(0, 1, 0) - x.in(0)) * x.M(x.m.divg_u));
    else x.out(0, (x.in(0, 0, 1) - x.in(0)) * x.M(x.m.divg_v));
  }
};
// divg = ((uc(i,j-1)-uc(i,j)) + (vc(i-1,j)-vc(i,j))) * rarea_c with the corner terms; levels
// that finished (nord < it) pass dd through.  in: ucw vcw ddprev ; out: dd
struct S_dd_div {
  static constexpr int NI = 3, NO = 1;
  struct P { LevOrd nord; int it; };
  static constexpr int NT = 5;
  static constexpr Tap taps[NT] = {{0, 0, -1, 0}, {0, 0, 0, 0}, {1, -1, 0, 0}, {1, 0, 0, 0}, {2, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    const int n = p.nord.v[x.kk];
    if (n < p.it) { if (n >= 1 && x.in_rect(g.is, g.ie + 1, g.js, g.je + 1)) x.out(0, x.in(2)); return; }
    const int nt = n - p.it;
    if (!x.in_rect(g.is - nt, g.ie + 1 + nt, g.js - nt, g.je + 1 + nt)) return;
    T um = x.in(0, 0, -1), u0 = x.in(0);
    T d = (um - u0) + (x.in(1, -1, 0) - x.in(1));
    if ((x.i == 1 || x.i == g.npx) && x.j == 1) d = d - um;
    if ((x.i == 1 || x.i == g.npx) && x.j == g.npy) d = d + u0;
    x.out(0, d * x.M(x.m.rarea_c));
  }
};

// damping term added to ke (:1341-1343 nord = 0, :1402-1432 nord > 0)
// in: delpc0 divg_d vq dd ; out: vd
struct S_ddamp {
  static constexpr int NI = 4, NO = 1;
  struct P { LevOrd nord; LevD d2_bg; double dddmp, d4_bg, dt; };
  static constexpr int NT = 4;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, 0, 0, 0}, {3, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie + 1, g.js, g.je + 1)) return;
    const int n = p.nord.v[x.kk];
    const double d2 = p.d2_bg.v[x.kk], dac = x.m.da_min_c;
    if (n == 0) {
      T dpc = x.in(0);
      T damp = dac * m_max(d2, m_min(0.20, p.dddmp * m_abs(dpc * p.dt)));
      x.out(0, damp * dpc);
    } else {
      T dpc = x.in(1);
      T vort;
      if (p.dddmp < 1.e-5) vort = T(0.0);
      else { T vq = x.in(2); vort = fabs(p.dt) * m_sqrt(dpc * dpc + vq * vq); }
      double dd8 = pow(dac * p.d4_bg, (double)(n + 1));
      T damp2 = dac * m_max(d2, m_min(0.20, p.dddmp * vort));
      x.out(0, damp2 * dpc + dd8 * x.in(3));
    }
  }
};

// absolute vorticity for the transport (:1453).  in: wk ; out: vort
struct S_absvort {
  static constexpr int NI = 1, NO = 1;
  struct P { int dummy; };
  static constexpr int NT = 1;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P&) {
    const Geom& g = x.g;
    if (!x.in_rect(g.is - g.ng, g.ie + g.ng, g.js - g.ng, g.je + g.ng)) return;
    x.out(0, x.in(0) + x.M(x.m.f0));
  }
};

// final D-grid wind update (:1474-1483, :1528-1539).  in: u v ke vd fxv fyv ut3 vt3 ; out: u_new v_new
struct S_duv {
  static constexpr int NI = 8, NO = 2;
  struct P { LevD vd_on; };
  static constexpr int NT = 12;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, 0, 0, 0}, {2, 1, 0, 0}, {2, 0, 1, 0},
                                   {3, 0, 0, 0}, {3, 1, 0, 0}, {3, 0, 1, 0}, {4, 0, 0, 0}, {5, 0, 0, 0}, {6, 0, 0, 0}, {7, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    const Geom& g = x.g;
    const double on = p.vd_on.v[x.kk];
    if (x.in_rect(g.is, g.ie, g.js, g.je + 1)) {
      auto ke0 = x.in(2) + x.in(3), ke1 = x.in(2, 1, 0) + x.in(3, 1, 0);
      x.out(0, x.in(0) * x.M(x.m.dx) + (ke0 - ke1) + x.in(5) + on * x.in(7));
    }
    if (x.in_rect(g.is, g.ie + 1, g.js, g.je)) {
      auto ke0 = x.in(2) + x.in(3), ke1 = x.in(2, 0, 1) + x.in(3, 0, 1);
      x.out(1, x.in(1) * x.M(x.m.dy) + (ke0 - ke1) - x.in(4) - on * x.in(6));
    }
  }
};

// ---- dissipative heating, d_con > 0 (:1436-1446, :1494-1525).  Edge quantities first:
//   ubh = (vd(i,j) - vd(i+1,j) + vt3) rdx , fyh = u' rdx   on (is:ie, js:je+1),  u' = the wind before the diffusive flux is added
//   vbh = (vd(i,j) - vd(i,j+1) - ut3) rdy , fxh = v' rdy   on (is:ie+1, js:je)
// in: u v ke vd fxv fyv ut3 vt3 (as S_duv) ; out: ubh fyh vbh fxh
struct S_dheat_edge {
  static constexpr int NI = 8, NO = 4;
  struct P { int dummy; };
  static constexpr int NT = 12;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, 0, 0, 0}, {2, 1, 0, 0}, {2, 0, 1, 0},
                                   {3, 0, 0, 0}, {3, 1, 0, 0}, {3, 0, 1, 0}, {4, 0, 0, 0}, {5, 0, 0, 0}, {6, 0, 0, 0}, {7, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P&) {
    const Geom& g = x.g;
    if (x.in_rect(g.is, g.ie, g.js, g.je + 1)) {
      auto ke0 = x.in(2) + x.in(3), ke1 = x.in(2, 1, 0) + x.in(3, 1, 0);
      const double rdx = x.M(x.m.rdx);
      x.out(0, ((x.in(3) - x.in(3, 1, 0)) + x.in(7)) * rdx);
      x.out(1, (x.in(0) * x.M(x.m.dx) + (ke0 - ke1) + x.in(5)) * rdx);
    }
    if (x.in_rect(g.is, g.ie + 1, g.js, g.je)) {
      auto ke0 = x.in(2) + x.in(3), ke1 = x.in(2, 0, 1) + x.in(3, 0, 1);
      const double rdy = x.M(x.m.rdy);
      x.out(2, ((x.in(3) - x.in(3, 0, 1)) - x.in(6)) * rdy);
      x.out(3, (x.in(1) * x.M(x.m.dy) + (ke0 - ke1) - x.in(4)) * rdy);
    }
  }
};
// heat_s of one level: the w-damping part (:938-951, ke_bg = 0) and, where d_con_k > 1e-5, the kinetic energy lost to the
// divergence / vorticity damping converted to heat (:1494-1525); levels with d_con_k = 0 (sponge layers, dyn_core_nlm.F90:595-628)
// keep the w part as it is -- not multiplied by delp, like the reference.
// in: delp_new ubh fyh vbh fxh w fx2w fy2w ; out: heat_s
struct S_dheat {
  static constexpr int NI = 8, NO = 1;
  struct P { int nonhydro; LevD dw_on, d_con; };
  static constexpr int NT = 14;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {1, 0, 1, 0}, {2, 0, 0, 0}, {2, 0, 1, 0}, {3, 0, 0, 0}, {3, 1, 0, 0},
                                   {4, 0, 0, 0}, {4, 1, 0, 0}, {5, 0, 0, 0}, {6, 0, 0, 0}, {6, 1, 0, 0}, {7, 0, 0, 0}, {7, 0, 1, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie, g.js, g.je)) return;
    T heat = T(0.0);
    if (p.nonhydro && p.dw_on.v[x.kk] != 0.0) {
      T dw = ((x.in(6) - x.in(6, 1, 0)) + (x.in(7) - x.in(7, 0, 1))) * x.M(x.m.rarea);
      heat = T(0.0) - dw * (x.in(5) + 0.5 * dw);
    }
    const double dcon = p.d_con.v[x.kk];
    if (dcon > 1.e-5) {
      const double damp = 0.25 * dcon;
      T ub0 = x.in(1), ub1 = x.in(1, 0, 1), fy0 = x.in(2), fy1 = x.in(2, 0, 1);
      T vb0 = x.in(3), vb1 = x.in(3, 1, 0), fx0 = x.in(4), fx1 = x.in(4, 1, 0);
      T u2 = fy0 + fy1, du2 = ub0 + ub1, v2 = fx0 + fx1, dv2 = vb0 + vb1;
      heat = x.in(0) * (heat - damp * x.M(x.m.rsin2) * ((ub0 * ub0 + ub1 * ub1 + vb0 * vb0 + vb1 * vb1) +
                                                        2.0 * (fy0 * ub0 + fy1 * ub1 + fx0 * vb0 + fx1 * vb1) -
                                                        x.M(x.m.cosa_s) * (u2 * dv2 + v2 * du2 + du2 * dv2)));
    }
    x.out(0, heat);
  }
};

struct DswParams {
  LevOrd hord_mt, hord_vt, hord_tm, hord_dp;
  LevOrd nord, nord_v, nord_w, nord_t;
  LevD d2_bg, damp_v, damp_w, damp_t;
  double dddmp, d4_bg, dt;
  bool hydrostatic;
  bool split_damp = false;   // perturbation side only: divergence damping evaluated separately for trajectory and perturbation
  bool heat = false;   // d_con > 1e-5: build the dissipative-heating stages
  LevD d_con;          // d_con_k per level (0 in the sponge layers)
};
struct DswOut { int delp, pt, u, v, w, fx, fy, crx, cry, xfx, yfx; int heat = -1; int divg = -1; };   // divg: the corner divergence d_sw leaves in its first argument (want_divg)
DswOut build_d_sw(Program& P, Mosaic& mo, int delp, int pt, int u, int v, int w, int uc, int vc, int ua, int va, int divg_d,
                  const DswParams& prm, int nk, const std::string& tag, const DswParams* pert = nullptr, bool want_divg = false);
// a2b_ord4 (a2b.cu)
int build_a2b_ord4(Program& P, Mosaic& mo, int qin, int nk, const std::string& tag);

}  // namespace fv3lm
