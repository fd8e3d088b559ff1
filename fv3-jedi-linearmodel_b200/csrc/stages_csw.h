// c_sw: C-grid half step (model/sw_core_nlm.F90:77-486; TL model_tlmadm/sw_core_tlm.F90:87,
// AD sw_core_adm.F90:112/704) with its helpers d2a2c_vect (:2746-3085),
// divergence_corner (:1661-1766) and fill_4corners (:3235).  Whole-tile edges/corners
// are keyed on the tile-global index exactly like the reference (i==1, i==npx, ...).
#pragma once
#include "engine.h"
#include "mosaic.h"

namespace fv3lm {
namespace csw {
constexpr double a1 = 0.5625, a2 = -0.0625;
constexpr double c1 = -2.0 / 14.0, c2 = 11.0 / 14.0, c3 = 5.0 / 14.0;

template <class T> DEV T edge_interpolate4(T u1, T u2, T u3, T u4, double d1, double d2, double d3, double d4) {
  double t1 = d1 + d2, t2 = d3 + d4;
  return 0.5 * (((t1 + d2) * u2 - d2 * u1) / t1 + ((t2 + d3) * u3 - d3 * u4) / t2);
}
}  // namespace csw

// (u, v) -> utmp, vtmp : D-grid winds averaged to cell centres (sw_core_nlm.F90:2815-2873)
struct S_d2a {
  static constexpr int NI = 2, NO = 2;
  struct P { int dummy; };
  static constexpr int NT = 8;
  static constexpr Tap taps[NT] = {{0, 0, -1, 0}, {0, 0, 0, 0}, {0, 0, 1, 0}, {0, 0, 2, 0},
                                   {1, -1, 0, 0}, {1, 0, 0, 0}, {1, 1, 0, 0}, {1, 2, 0, 0}};
  template <class X> DEV static void eval(X& x, const P&) {
    const Geom& g = x.g;
    const int isd = g.is - g.ng, ied = g.ie + g.ng, jsd = g.js - g.ng, jed = g.je + g.ng;
    if (!x.in_rect(isd, ied, jsd, jed)) return;
    const int npt = 4;
    const bool inner = x.i >= npt && x.i <= g.npx - npt && x.j >= npt && x.j <= g.npy - npt;
    // 4th-order interior formula only on the rows / columns the reference computes it on
    // (utmp: j = max(npt,js-1)..min(npy-npt,je+1); vtmp: i = max(npt,is-1)..min(npx-npt,ie+1), sw_core_nlm.F90:2815-2826);
    // on a sub-domain the remaining halo rows have no neighbours two cells out and are never consumed
    const bool inner_u = inner && x.jl >= g.js - 1 && x.jl <= g.je + 1;
    const bool inner_v = inner && x.il >= g.is - 1 && x.il <= g.ie + 1;
    if (inner_u) x.out(0, csw::a2 * (x.in(0, 0, -1) + x.in(0, 0, 2)) + csw::a1 * (x.in(0, 0, 0) + x.in(0, 0, 1)));
    else x.out(0, 0.5 * (x.in(0, 0, 0) + x.in(0, 0, 1)));
    if (inner_v) x.out(1, csw::a2 * (x.in(1, -1, 0) + x.in(1, 2, 0)) + csw::a1 * (x.in(1, 0, 0) + x.in(1, 1, 0)));
    else x.out(1, 0.5 * (x.in(1, 0, 0) + x.in(1, 1, 0)));
  }
};

// (utmp, vtmp) -> ua, va  contravariant A-grid winds (:2875-2880)
struct S_uava {
  static constexpr int NI = 2, NO = 2;
  struct P { int dummy; };
  static constexpr int NT = 2;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P&) {
    const Geom& g = x.g;
    if (!x.in_rect(g.is - 2, g.ie + 2, g.js - 2, g.je + 2)) return;
    double cs = x.M(x.m.cosa_s), r2 = x.M(x.m.rsin2);
    auto ut = x.in(0), vt = x.in(1);
    x.out(0, (ut - vt * cs) * r2);
    x.out(1, (vt - ut * cs) * r2);
  }
};

// A -> C in x:  (utmp, v, ua) -> uc, ut(scaled by dt2*dy*sin)      (:2904-2984, c_sw :157-165)
// DIR = 1 is the y analogue (vtmp, u, va) -> vc, vt                 (:3034-3082, c_sw :166-174)
template <int DIR> struct S_a2c {
  static constexpr int NI = 3, NO = 2;
  struct P { double dt2; };
  static constexpr int NT = 9;
  static constexpr Tap taps[NT] = {
      {0, DIR == 0 ? -2 : 0, DIR == 0 ? 0 : -2, 0}, {0, DIR == 0 ? -1 : 0, DIR == 0 ? 0 : -1, 0}, {0, 0, 0, 0},
      {0, DIR == 0 ? 1 : 0, DIR == 0 ? 0 : 1, 0},   {1, 0, 0, 0},
      {2, DIR == 0 ? -2 : 0, DIR == 0 ? 0 : -2, 0}, {2, DIR == 0 ? -1 : 0, DIR == 0 ? 0 : -1, 0}, {2, 0, 0, 0},
      {2, DIR == 0 ? 1 : 0, DIR == 0 ? 0 : 1, 0}};
  template <class X> DEV static typename X::T A(const X& x, int f, int d) { return DIR == 0 ? x.in(f, d, 0) : x.in(f, 0, d); }
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (DIR == 0) { if (!x.in_rect(g.is - 1, g.ie + 2, g.js - 1, g.je + 1)) return; }
    else { if (!x.in_rect(g.is - 1, g.ie + 1, g.js - 1, g.je + 2)) return; }
    const int pos = DIR == 0 ? x.i : x.j;
    const int np = DIR == 0 ? g.npx : g.npy;
    const double* cosa = DIR == 0 ? x.m.cosa_u : x.m.cosa_v;
    const double* rsin = DIR == 0 ? x.m.rsin_u : x.m.rsin_v;
    const double* da = DIR == 0 ? x.m.dxa : x.m.dya;
    const double* s_lo = DIR == 0 ? x.m.sin_sg3 : x.m.sin_sg4;   // sin_sg(i-1,j,3) / sin_sg(i,j-1,4)
    const double* s_hi = DIR == 0 ? x.m.sin_sg1 : x.m.sin_sg2;   // sin_sg(i,j,1)   / sin_sg(i,j,2)
    auto MD = [&](const double* a, int d) { return DIR == 0 ? x.M(a, d, 0) : x.M(a, 0, d); };
    T c, t;
    if (pos == 1 || pos == np) {
      t = csw::edge_interpolate4(A(x, 2, -2), A(x, 2, -1), A(x, 2, 0), A(x, 2, 1), MD(da, -2), MD(da, -1), MD(da, 0), MD(da, 1));
      c = val(t) > 0.0 ? t * MD(s_lo, -1) : t * MD(s_hi, 0);
    } else {
      if (pos == 0 || pos == np - 1) c = csw::c1 * A(x, 0, -2) + csw::c2 * A(x, 0, -1) + csw::c3 * A(x, 0, 0);
      else if (pos == 2 || pos == np + 1) c = csw::c1 * A(x, 0, 1) + csw::c2 * A(x, 0, 0) + csw::c3 * A(x, 0, -1);
      else c = csw::a2 * (A(x, 0, -2) + A(x, 0, 1)) + csw::a1 * (A(x, 0, -1) + A(x, 0, 0));
      t = (c - x.in(1) * x.M(cosa)) * x.M(rsin);
    }
    x.out(0, c);
    const double* dd = DIR == 0 ? x.m.dy : x.m.dx;
    T ts = val(t) > 0.0 ? p.dt2 * t * x.M(dd) * MD(s_lo, -1) : p.dt2 * t * x.M(dd) * MD(s_hi, 0);
    x.out(1, ts);
  }
};

// divergence_corner (u, v, ua, va) -> divg_d   (:1722-1763)
struct S_divg_corner {
  static constexpr int NI = 4, NO = 1;
  struct P { int dummy; };
  static constexpr int NT = 12;
  static constexpr Tap taps[NT] = {{0, -1, 0, 0}, {0, 0, 0, 0}, {1, 0, -1, 0}, {1, 0, 0, 0},
                                   {2, -1, -1, 0}, {2, 0, -1, 0}, {2, -1, 0, 0}, {2, 0, 0, 0},
                                   {3, -1, -1, 0}, {3, -1, 0, 0}, {3, 0, -1, 0}, {3, 0, 0, 0}};
  template <class X> DEV static typename X::T uf(const X& x, int di) {   // uf(i+di, j)
    double s = 0.5 * (x.M(x.m.sin_sg4, di, -1) + x.M(x.m.sin_sg2, di, 0));
    if (x.j == 1 || x.j == x.g.npy) return x.in(0, di, 0) * x.M(x.m.dyc, di, 0) * s;
    return (x.in(0, di, 0) - 0.25 * (x.in(3, di, -1) + x.in(3, di, 0)) * (x.M(x.m.cos_sg4, di, -1) + x.M(x.m.cos_sg2, di, 0))) *
           x.M(x.m.dyc, di, 0) * s;
  }
  template <class X> DEV static typename X::T vf(const X& x, int dj) {   // vf(i, j+dj)
    double s = 0.5 * (x.M(x.m.sin_sg3, -1, dj) + x.M(x.m.sin_sg1, 0, dj));
    if (x.i == 1 || x.i == x.g.npx) return x.in(1, 0, dj) * x.M(x.m.dxc, 0, dj) * s;
    return (x.in(1, 0, dj) - 0.25 * (x.in(2, -1, dj) + x.in(2, 0, dj)) * (x.M(x.m.cos_sg3, -1, dj) + x.M(x.m.cos_sg1, 0, dj))) *
           x.M(x.m.dxc, 0, dj) * s;
  }
  template <class X> DEV static void eval(X& x, const P&) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie + 1, g.js, g.je + 1)) return;
    T vm = vf(x, -1), v0 = vf(x, 0);
    T d = (vm - v0) + (uf(x, -1) - uf(x, 0));
    if ((x.i == 1 || x.i == g.npx) && x.j == 1) d = d - vm;
    if ((x.i == 1 || x.i == g.npx) && x.j == g.npy) d = d + v0;
    x.out(0, x.M(x.m.rarea_c) * d);
  }
};

// first-order upwind fluxes of delp, pt, w  (c_sw :196-228 x, :233-274 y)
// in: 0 = ut/vt (scaled), 1 = delp, 2 = pt, 3 = w ; out: fx1, fx, fx2
template <int DIR> struct S_cflux {
  static constexpr int NI = 4, NO = 3;
  struct P { int nonhydro; };
  static constexpr int NT = 7;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, DIR == 0 ? -1 : 0, DIR == 0 ? 0 : -1, 0}, {1, 0, 0, 0},
                                   {2, DIR == 0 ? -1 : 0, DIR == 0 ? 0 : -1, 0}, {2, 0, 0, 0},
                                   {3, DIR == 0 ? -1 : 0, DIR == 0 ? 0 : -1, 0}, {3, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (DIR == 0) { if (!x.in_rect(g.is - 1, g.ie + 2, g.js - 1, g.je + 1)) return; }
    else { if (!x.in_rect(g.is - 1, g.ie + 1, g.js - 1, g.je + 2)) return; }
    T c = x.in(0);
    const int di = (DIR == 0 && val(c) > 0.0) ? -1 : 0, dj = (DIR == 1 && val(c) > 0.0) ? -1 : 0;
    T f1 = c * x.in(1, di, dj);
    x.out(0, f1);
    x.out(1, f1 * x.in(2, di, dj));
    if (p.nonhydro) x.out(2, f1 * x.in(3, di, dj));
  }
};

// delpc, ptc, wc  (c_sw :246-283)
// in: delp pt w fx1 fx fx2 fy1 fy fy2 ; out: delpc ptc wc
struct S_cupd {
  static constexpr int NI = 9, NO = 3;
  struct P { int nonhydro; };
  static constexpr int NT = 15;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, 0, 0, 0},
                                   {3, 0, 0, 0}, {3, 1, 0, 0}, {4, 0, 0, 0}, {4, 1, 0, 0}, {5, 0, 0, 0}, {5, 1, 0, 0},
                                   {6, 0, 0, 0}, {6, 0, 1, 0}, {7, 0, 0, 0}, {7, 0, 1, 0}, {8, 0, 0, 0}, {8, 0, 1, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is - 1, g.ie + 1, g.js - 1, g.je + 1)) return;
    double ra = x.M(x.m.rarea);
    T dp = x.in(0);
    T dpc = dp + ((x.in(3) - x.in(3, 1, 0)) + (x.in(6) - x.in(6, 0, 1))) * ra;
    x.out(0, dpc);
    x.out(1, (x.in(1) * dp + ((x.in(4) - x.in(4, 1, 0)) + (x.in(7) - x.in(7, 0, 1))) * ra) / dpc);
    if (p.nonhydro) x.out(2, (x.in(2) * dp + ((x.in(5) - x.in(5, 1, 0)) + (x.in(8) - x.in(8, 0, 1))) * ra) / dpc);
  }
};

// kinetic energy on the A grid (c_sw :314-364).  in: ua va uc vc u v ; out: ke
struct S_cke {
  static constexpr int NI = 6, NO = 1;
  struct P { double dt2; };
  static constexpr int NT = 10;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, 0, 0, 0}, {2, 1, 0, 0}, {3, 0, 0, 0}, {3, 0, 1, 0},
                                   {4, 0, 0, 0}, {4, 0, 1, 0}, {5, 0, 0, 0}, {5, 1, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is - 1, g.ie + 1, g.js - 1, g.je + 1)) return;
    T ua = x.in(0), va = x.in(1), ke, vo;
    if (val(ua) > 0.0) {
      if (x.i == 1 || x.i == g.npx) ke = x.in(2) * x.M(x.m.sin_sg1) + x.in(5) * x.M(x.m.cos_sg1);
      else ke = x.in(2);
    } else {
      if (x.i == 0 || x.i == g.npx - 1) ke = x.in(2, 1, 0) * x.M(x.m.sin_sg3) + x.in(5, 1, 0) * x.M(x.m.cos_sg3);
      else ke = x.in(2, 1, 0);
    }
    if (val(va) > 0.0) {
      if (x.j == 1 || x.j == g.npy) vo = x.in(3) * x.M(x.m.sin_sg2) + x.in(4) * x.M(x.m.cos_sg2);
      else vo = x.in(3);
    } else {
      if (x.j == 0 || x.j == g.npy - 1) vo = x.in(3, 0, 1) * x.M(x.m.sin_sg4) + x.in(4, 0, 1) * x.M(x.m.cos_sg4);
      else vo = x.in(3, 0, 1);
    }
    x.out(0, 0.5 * p.dt2 * (ua * ke + va * vo));
  }
};

// absolute vorticity at cell corners (c_sw :370-401).  in: uc vc ; out: vort
struct S_cvort {
  static constexpr int NI = 2, NO = 1;
  struct P { int dummy; };
  static constexpr int NT = 4;
  static constexpr Tap taps[NT] = {{0, 0, -1, 0}, {0, 0, 0, 0}, {1, 0, 0, 0}, {1, -1, 0, 0}};
  template <class X> DEV static void eval(X& x, const P&) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie + 1, g.js, g.je + 1)) return;
    T fxm = x.in(0, 0, -1) * x.M(x.m.dxc, 0, -1), fx0 = x.in(0) * x.M(x.m.dxc);
    T fy0 = x.in(1) * x.M(x.m.dyc), fym = x.in(1, -1, 0) * x.M(x.m.dyc, -1, 0);
    T v = (fxm - fx0) + (fy0 - fym);
    if (x.i == 1 && (x.j == 1 || x.j == g.npy)) v = v + fym;
    if (x.i == g.npx && (x.j == 1 || x.j == g.npy)) v = v - fy0;
    x.out(0, x.M(x.m.fC) + x.M(x.m.rarea_c) * v);
  }
};

// time-centred C-grid winds (c_sw :434-484).  in: uc vc u v vort ke ; out: uc_new vc_new
struct S_cwind {
  static constexpr int NI = 6, NO = 2;
  struct P { double dt2; };
  static constexpr int NT = 10;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, 0, 0, 0}, {3, 0, 0, 0},
                                   {4, 0, 0, 0}, {4, 0, 1, 0}, {4, 1, 0, 0},
                                   {5, -1, 0, 0}, {5, 0, 0, 0}, {5, 0, -1, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (x.in_rect(g.is, g.ie + 1, g.js, g.je)) {
      T fy1;
      if (x.i == 1 || x.i == g.npx) fy1 = p.dt2 * x.in(3);
      else fy1 = p.dt2 * (x.in(3) - x.in(0) * x.M(x.m.cosa_u)) / x.M(x.m.sina_u);
      T fy = val(fy1) > 0.0 ? x.in(4) : x.in(4, 0, 1);
      x.out(0, x.in(0) + fy1 * fy + x.M(x.m.rdxc) * (x.in(5, -1, 0) - x.in(5)));
    }
    if (x.in_rect(g.is, g.ie, g.js, g.je + 1)) {
      T fx1;
      if (x.j == 1 || x.j == g.npy) fx1 = p.dt2 * x.in(2);
      else fx1 = p.dt2 * (x.in(2) - x.in(1) * x.M(x.m.cosa_v)) / x.M(x.m.sina_v);
      T fx = val(fx1) > 0.0 ? x.in(4) : x.in(4, 1, 0);
      x.out(1, x.in(1) - fx1 * fx + x.M(x.m.rdyc) * (x.in(5, 0, -1) - x.in(5)));
    }
  }
};

struct CswOut { int delpc, ptc, wc, uc, vc, ua, va, ut, vt, divg_d; };
CswOut build_c_sw(Program& P, Mosaic& mo, int delp, int pt, int u, int v, int w, double dt2, bool hydrostatic, int nord,
                  int nk, const std::string& tag);

}  // namespace fv3lm
