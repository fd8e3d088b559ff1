// a2b_ord4: 4th-order A-grid -> B-grid interpolation (model/a2b_edge_nlm.F90:49-329,
// extrap_corner :800-810; TL a2b_edge_tlm.F90:546, AD a2b_edge_adm.F90:72).  Linear in qin.
#include "stages_dsw.h"
#include "modules.h"
#include "fused_tp.h"

namespace fv3lm {
namespace a2b {
constexpr double a1 = 0.5625, a2 = -0.0625, b1 = 7.0 / 12.0, b2 = -1.0 / 12.0, c1 = 2.0 / 3.0, c2 = -1.0 / 6.0, r3 = 1.0 / 3.0;

}  // namespace a2b

// qx (DIR 0) / qy (DIR 1): 1-D PPM-form interpolation to cell faces (:108-149 / :152-194)
template <int DIR> struct S_a2b_q1 {
  static constexpr int NI = 1, NO = 1;
  struct P { int dummy; };
  static constexpr int NT = 6;
  static constexpr Tap taps[NT] = {{0, DIR == 0 ? -3 : 0, DIR == 0 ? 0 : -3, 0}, {0, DIR == 0 ? -2 : 0, DIR == 0 ? 0 : -2, 0},
                                   {0, DIR == 0 ? -1 : 0, DIR == 0 ? 0 : -1, 0}, {0, 0, 0, 0},
                                   {0, DIR == 0 ? 1 : 0, DIR == 0 ? 0 : 1, 0},   {0, DIR == 0 ? 2 : 0, DIR == 0 ? 0 : 2, 0}};
  template <class X> DEV static typename X::T Q(const X& x, int d) { return DIR == 0 ? x.in(0, d, 0) : x.in(0, 0, d); }
  template <class X> DEV static double DA(const X& x, int d) { return DIR == 0 ? x.M(x.m.dxa, d, 0) : x.M(x.m.dya, 0, d); }
  template <class X> DEV static typename X::T gen(const X& x, int d) {   // generic value at face pos+d
    return a2b::b2 * (Q(x, d - 2) + Q(x, d + 1)) + a2b::b1 * (Q(x, d - 1) + Q(x, d));
  }
  // value at the tile-edge face (index 1 or np), located at offset d from the current point
  template <class X> DEV static typename X::T edge(const X& x, int d, bool low) {
    // low edge: inside cell is (pos+d), outside (pos+d-1); high edge: inside (pos+d-1), outside (pos+d)
    if (low) {
      double g_in = DA(x, d + 1) / DA(x, d), g_ou = DA(x, d - 2) / DA(x, d - 1);
      return 0.5 * (((2.0 + g_in) * Q(x, d) - Q(x, d + 1)) / (1.0 + g_in) + ((2.0 + g_ou) * Q(x, d - 1) - Q(x, d - 2)) / (1.0 + g_ou));
    }
    double g_in = DA(x, d - 2) / DA(x, d - 1), g_ou = DA(x, d + 1) / DA(x, d);
    return 0.5 * (((2.0 + g_in) * Q(x, d - 1) - Q(x, d - 2)) / (1.0 + g_in) + ((2.0 + g_ou) * Q(x, d) - Q(x, d + 1)) / (1.0 + g_ou));
  }
  template <class X> DEV static void eval(X& x, const P&) {
    using T = typename X::T;
    const Geom& g = x.g;
    const int np = DIR == 0 ? g.npx : g.npy;
    // a2b_edge_nlm.F90:108-110 / :152-154: rows js-2..je+2 (columns is-2..ie+2) clipped to the tile
    if (DIR == 0) { if (!x.in_rect(g.is, g.ie + 1, g.js - 2, g.je + 2) || !x.in_tile(1, g.npx, 1, g.npy - 1)) return; }
    else { if (!x.in_rect(g.is - 2, g.ie + 2, g.js, g.je + 1) || !x.in_tile(1, g.npx - 1, 1, g.npy)) return; }
    const int pos = DIR == 0 ? x.i : x.j;
    T r;
    if (pos == 1) r = edge(x, 0, true);
    else if (pos == np) r = edge(x, 0, false);
    else if (pos == 2) {
      double g_in = DA(x, 0) / DA(x, -1);
      r = (3.0 * (g_in * Q(x, -1) + Q(x, 0)) - (g_in * edge(x, -1, true) + gen(x, 1))) / (2.0 + 2.0 * g_in);
    } else if (pos == np - 1) {
      double g_in = DA(x, -1) / DA(x, 0);
      r = (3.0 * (Q(x, -1) + g_in * Q(x, 0)) - (g_in * edge(x, 1, false) + gen(x, -1))) / (2.0 + 2.0 * g_in);
    } else r = gen(x, 0);
    x.out(0, r);
  }
  // gather adjoint: away from the tile edges the operator is the uniform 4-point stencil b2, b1, b1, b2 --
  // four loads instead of six seeded evaluations; near the edges the generic path is used
  static constexpr bool custom_ad = true;
  template <class K> DEV static void adjoint(const K& kn, int ii, int jj, int kk, int tile, double* acc) {
    if (kk >= kn.nk_fwd || !kn.outad.p[0]) return;
    const Geom& g = kn.g;
    CtxNL<S_a2b_q1<DIR>> x; x.g = kn.g; x.m = kn.m; x.in_ = kn.in;
    x.setpos(ii, jj, kk, tile, g.i0[tile], g.j0[tile]);
    const int pos = DIR == 0 ? x.i : x.j, np = DIR == 0 ? g.npx : g.npy;
    const int loc = DIR == 0 ? x.il : x.jl, cl = DIR == 0 ? x.jl : x.il, cg = DIR == 0 ? x.j : x.i;
    const int lo = DIR == 0 ? g.is : g.js, hi = (DIR == 0 ? g.ie : g.je) + 1;          // output range along DIR
    const int clo = (DIR == 0 ? g.js : g.is) - 2, chi = (DIR == 0 ? g.je : g.ie) + 2;  // across DIR
    const int ncross = DIR == 0 ? g.npy : g.npx;
    // (the faces 2 and np-1 also read the generic value of their inner neighbour: stay one more cell away)
    if (pos - 1 >= 4 && pos + 2 <= np - 3 && loc - 1 >= lo && loc + 2 <= hi && cl >= clo && cl <= chi && cg >= 1 && cg <= ncross - 1) {
      const int stride = DIR == 0 ? 1 : g.pitch;
      const double* ap = kn.outad.p[0] + x.off(kn.outad.nk[0], 0, 0, 0);
      acc[0] += a2b::b2 * (LDG(ap - stride) + LDG(ap + 2 * stride)) + a2b::b1 * (LDG(ap) + LDG(ap + stride));
    } else {
      AdTaps<S_a2b_q1<DIR>, 0>::run(kn, ii, jj, kk, tile, acc);
    }
  }
};

// qout on the tile boundary: 3-way corner extrapolation (:73-106) and edge interpolation
// with edge_w/e/s/n (:116-121, :134-139, :159-164, :178-183).   in: qin ; out: qe
struct S_a2b_edge {
  static constexpr int NI = 1, NO = 1;
  struct P { int dummy; };
  static constexpr int NT = 8;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {0, 1, 1, 0}, {0, -1, 0, 0}, {0, -2, 1, 0}, {0, 0, -1, 0}, {0, 1, -2, 0}, {0, -1, -1, 0}, {0, -2, -2, 0}};
  // extrap_corner(p0 = this corner, p1 = agrid(a1), p2 = agrid(a2), q1, q2) = q1 + x1/(x2-x1) (q1-q2) with the
  // great-circle distances x1, x2 (a2b_edge_nlm.F90:800-810).  The weight is pure geometry: it is precomputed once
  // per tile corner (fv3lm_set_metric -> a2b_cw[corner * 3 + n]) instead of 6 haversines per corner thread.
  template <class X> DEV static typename X::T corner(const X& x, int cn, int a1i, int a1j, int a2i, int a2j) {
    const double w = x.m.a2b_cw[(size_t)x.tile * x.g.slab + cn];
    auto q1 = x.in(0, a1i - x.i, a1j - x.j), q2 = x.in(0, a2i - x.i, a2j - x.j);
    return q1 + w * (q1 - q2);
  }
  template <class X> DEV static void eval(X& x, const P&) {
    using T = typename X::T;
    const Geom& g = x.g;
    const int i = x.i, j = x.j, npx = g.npx, npy = g.npy;
    if (!x.in_rect(g.is, g.ie + 1, g.js, g.je + 1)) return;
    const bool ei = (i == 1 || i == npx), ej = (j == 1 || j == npy);
    if (!ei && !ej) return;
    T r;
    if (ei && ej) {
      if (i == 1 && j == 1) r = (corner(x, 0, 1, 1, 2, 2) + corner(x, 1, 0, 1, -1, 2) + corner(x, 2, 1, 0, 2, -1)) * a2b::r3;
      else if (i == npx && j == 1) r = (corner(x, 3, npx - 1, 1, npx - 2, 2) + corner(x, 4, npx - 1, 0, npx - 2, -1) + corner(x, 5, npx, 1, npx + 1, 2)) * a2b::r3;
      else if (i == npx && j == npy) r = (corner(x, 6, npx - 1, npy - 1, npx - 2, npy - 2) + corner(x, 7, npx, npy - 1, npx + 1, npy - 2) + corner(x, 8, npx - 1, npy, npx - 2, npy + 1)) * a2b::r3;
      else r = (corner(x, 9, 1, npy - 1, 2, npy - 2) + corner(x, 10, 0, npy - 1, -1, npy - 2) + corner(x, 11, 1, npy, 2, npy + 1)) * a2b::r3;
    } else if (ei) {
      // q2(j) = (qin(i-1,j)*dxa(i,j) + qin(i,j)*dxa(i-1,j)) / (dxa(i-1,j)+dxa(i,j))
      auto q2 = [&](int dj) { return (x.in(0, -1, dj) * x.M(x.m.dxa, 0, dj) + x.in(0, 0, dj) * x.M(x.m.dxa, -1, dj)) / (x.M(x.m.dxa, -1, dj) + x.M(x.m.dxa, 0, dj)); };
      double e = x.M1(i == 1 ? x.m.edge_w : x.m.edge_e, j);
      r = e * q2(-1) + (1.0 - e) * q2(0);
    } else {
      auto q1 = [&](int di) { return (x.in(0, di, -1) * x.M(x.m.dya, di, 0) + x.in(0, di, 0) * x.M(x.m.dya, di, -1)) / (x.M(x.m.dya, di, -1) + x.M(x.m.dya, di, 0)); };
      double e = x.M1(j == 1 ? x.m.edge_s : x.m.edge_n, i);
      r = e * q1(-1) + (1.0 - e) * q1(0);
    }
    x.out(0, r);
  }
  // gather adjoint: only outputs on the tile boundary exist and they read at most two cells inwards
  static constexpr bool custom_ad = true;
  template <class K> DEV static void adjoint(const K& kn, int ii, int jj, int kk, int tile, double* acc) {
    const Geom& g = kn.g;
    const int i = ii - (g.ng - 1) + g.i0[tile], j = jj - (g.ng - 1) + g.j0[tile];
    if (i > 3 && i < g.npx - 3 && j > 3 && j < g.npy - 3) return;
    AdTaps<S_a2b_edge, 0>::run(kn, ii, jj, kk, tile, acc);
  }
};

// second sweep and averaging (:196-227).  in: qx qy qe ; out: qout
struct S_a2b_q2 {
  static constexpr int NI = 3, NO = 1;
  struct P { int dummy; };
  static constexpr int NT = 17;
  static constexpr Tap taps[NT] = {{0, 0, -3, 0}, {0, 0, -2, 0}, {0, 0, -1, 0}, {0, 0, 0, 0}, {0, 0, 1, 0}, {0, 0, 2, 0},
                                   {1, -3, 0, 0}, {1, -2, 0, 0}, {1, -1, 0, 0}, {1, 0, 0, 0}, {1, 1, 0, 0}, {1, 2, 0, 0},
                                   {2, 0, 0, 0}, {2, 0, -1, 0}, {2, 0, 1, 0}, {2, -1, 0, 0}, {2, 1, 0, 0}};
  template <class X> DEV static typename X::T xx(const X& x, int dj) {   // generic qxx(i, j+dj)
    return a2b::a2 * (x.in(0, 0, dj - 2) + x.in(0, 0, dj + 1)) + a2b::a1 * (x.in(0, 0, dj - 1) + x.in(0, 0, dj));
  }
  template <class X> DEV static typename X::T yy(const X& x, int di) {
    return a2b::a2 * (x.in(1, di - 2, 0) + x.in(1, di + 1, 0)) + a2b::a1 * (x.in(1, di - 1, 0) + x.in(1, di, 0));
  }
  template <class X> DEV static void eval(X& x, const P&) {
    using T = typename X::T;
    const Geom& g = x.g;
    const int i = x.i, j = x.j, npx = g.npx, npy = g.npy;
    if (!x.in_rect(g.is, g.ie + 1, g.js, g.je + 1)) return;
    if (i == 1 || i == npx || j == 1 || j == npy) { x.out(0, x.in(2)); return; }
    T qxx, qyy;
    if (j == 2) qxx = a2b::c1 * (x.in(0, 0, -1) + x.in(0)) + a2b::c2 * (x.in(2, 0, -1) + xx(x, 1));
    else if (j == npy - 1) qxx = a2b::c1 * (x.in(0, 0, -1) + x.in(0)) + a2b::c2 * (x.in(2, 0, 1) + xx(x, -1));
    else qxx = xx(x, 0);
    if (i == 2) qyy = a2b::c1 * (x.in(1, -1, 0) + x.in(1)) + a2b::c2 * (x.in(2, -1, 0) + yy(x, 1));
    else if (i == npx - 1) qyy = a2b::c1 * (x.in(1, -1, 0) + x.in(1)) + a2b::c2 * (x.in(2, 1, 0) + yy(x, -1));
    else qyy = yy(x, 0);
    x.out(0, 0.5 * (qxx + qyy));
  }
  // gather adjoint: five or more cells away from the tile edges the operator is 0.5 * (a2, a1, a1, a2) along j for qx
  // and along i for qy, and the edge values qe are not read at all
  static constexpr bool custom_ad = true;
  template <class K> DEV static void adjoint(const K& kn, int ii, int jj, int kk, int tile, double* acc) {
    if (kk >= kn.nk_fwd || !kn.outad.p[0]) return;
    const Geom& g = kn.g;
    CtxNL<S_a2b_q2> x; x.g = kn.g; x.m = kn.m; x.in_ = kn.in;
    x.setpos(ii, jj, kk, tile, g.i0[tile], g.j0[tile]);
    if (x.i >= 5 && x.i <= g.npx - 5 && x.j >= 5 && x.j <= g.npy - 5 &&
        x.il - 1 >= g.is && x.il + 2 <= g.ie + 1 && x.jl - 1 >= g.js && x.jl + 2 <= g.je + 1) {
      const double* ap = kn.outad.p[0] + x.off(kn.outad.nk[0], 0, 0, 0);
      const int st = g.pitch;
      acc[0] += 0.5 * (a2b::a2 * (LDG(ap - st) + LDG(ap + 2 * st)) + a2b::a1 * (LDG(ap) + LDG(ap + st)));
      acc[1] += 0.5 * (a2b::a2 * (LDG(ap - 1) + LDG(ap + 2)) + a2b::a1 * (LDG(ap) + LDG(ap + 1)));
    } else {
      AdTaps<S_a2b_q2, 0>::run(kn, ii, jj, kk, tile, acc);
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------------
// a2b_ord4 as ONE tile kernel per sweep direction (FV3LM_FUSED_A2B=1; opt-in until timed on a B200): qin -> qout with qx, qy and
// the boundary values qe in shared memory (2 array passes instead of 10).  The stages' own eval() run on tile contexts
// (fused_tp.h, TCtx).  The operator is linear in qin with purely geometric coefficients, so the reverse kernel is the transposed
// gather: qout_ad tile -> qx_ad, qy_ad, qe_ad tiles -> qin_ad of the block's own cells, with the uniform-stencil fast paths of
// the stage adjoints above away from the tile edges and seeded evaluations of the stages near them.
// ---------------------------------------------------------------------------------------------------------------------
namespace ftp {
template <class TT> struct KernA2b {
  static constexpr bool TLM = !std::is_same<TT, double>::value;
  static constexpr int NPH = 3;
  Geom g; Metrics m; int nk; Fld qin; OFld qout;
  struct Smem { SBuf<TLM, QW * QH> q; SBuf<TLM, TX * QH> qx; SBuf<TLM, QW * TY> qy; SBuf<TLM, (TX + 2) * (TY + 2)> qe; };
  DEV void phase(int ph, int tid, int bx, int by, int z, Smem& s) const {
    using N = Num<TT>;
    int tile, kk; split_z(z, nk, tile, kk);
    const int ii0 = bx * TX, jj0 = by * TY;
    const int i0 = g.i0[tile], j0 = g.j0[tile];
    const Box bq{ii0 - 3, jj0 - 3, QW, QH}, bqx{ii0, jj0 - 3, TX, QH}, bqy{ii0 - 3, jj0, QW, TY}, bqe{ii0 - 1, jj0 - 1, TX + 2, TY + 2};
    auto inside = [&](int ii, int jj) { return ii >= 0 && ii < g.NX && jj >= 0 && jj < g.NY; };
    if (ph == 0) {
      CtxBase x; x.g = g;
      for (int c = tid; c < bq.n(); c += NTHR) {
        const int ii = bq.x0 + c % bq.w, jj = bq.y0 + c / bq.w;
        TT a = TT(0.0);
        if (inside(ii, jj)) { x.setpos(ii, jj, kk, tile, i0, j0); a = N::ld(qin, x.off(qin.nk, 0, 0, 0)); }
        N::sts(s.q.v, s.q.d, c, a);
      }
    } else if (ph == 1) {
      TCtx<TT, 1, 1> x; x.g = g; x.m = m;
      x.ti[0] = TRef{s.q.v, s.q.d, bq}; x.gi[0] = Fld{nullptr, nullptr, 1}; x.go[0] = OFld{nullptr, nullptr, 1};
      x.to[0] = TOut{s.qx.v, s.qx.d, bqx};
      for (int c = tid; c < bqx.n(); c += NTHR) {
        const int ii = bqx.x0 + c % bqx.w, jj = bqx.y0 + c / bqx.w;
        if (!inside(ii, jj)) continue;
        x.setpos(ii, jj, kk, tile, i0, j0);
        S_a2b_q1<0>::eval(x, {0});
      }
      x.to[0] = TOut{s.qy.v, s.qy.d, bqy};
      for (int c = tid; c < bqy.n(); c += NTHR) {
        const int ii = bqy.x0 + c % bqy.w, jj = bqy.y0 + c / bqy.w;
        if (!inside(ii, jj)) continue;
        x.setpos(ii, jj, kk, tile, i0, j0);
        S_a2b_q1<1>::eval(x, {0});
      }
      x.to[0] = TOut{s.qe.v, s.qe.d, bqe};
      for (int c = tid; c < bqe.n(); c += NTHR) {
        const int ii = bqe.x0 + c % bqe.w, jj = bqe.y0 + c / bqe.w;
        if (!inside(ii, jj)) continue;
        x.setpos(ii, jj, kk, tile, i0, j0);
        S_a2b_edge::eval(x, {0});
      }
    } else {
      TCtx<TT, 3, 1> x; x.g = g; x.m = m;
      x.ti[0] = TRef{s.qx.v, s.qx.d, bqx}; x.ti[1] = TRef{s.qy.v, s.qy.d, bqy}; x.ti[2] = TRef{s.qe.v, s.qe.d, bqe};
      for (int f = 0; f < 3; f++) x.gi[f] = Fld{nullptr, nullptr, 1};
      x.to[0] = TOut{nullptr, nullptr, bq}; x.go[0] = qout;
      for (int c = tid; c < TX * TY; c += NTHR) {
        const int ii = ii0 + c % TX, jj = jj0 + c / TX;
        if (!inside(ii, jj)) continue;
        x.setpos(ii, jj, kk, tile, i0, j0);
        S_a2b_q2::eval(x, {0});
      }
    }
  }
};

struct KernA2bRev {
  static constexpr int NPH = 3;
  static constexpr int UW = TX + 5, UH = TY + 5;
  Geom g; Metrics m; int nk; Fld aout; OFld qin_ad;      // .v = adjoint arrays
  struct Smem { double oad[UW * UH], qx_ad[UW * TY], qy_ad[TX * UH], qe_ad[(TX + 3) * (TY + 3)]; };
  DEV void phase(int ph, int tid, int bx, int by, int z, Smem& s) const {
    int tile, kk; split_z(z, nk, tile, kk);
    const int ii0 = bx * TX, jj0 = by * TY;
    const int i0 = g.i0[tile], j0 = g.j0[tile];
    const Box bu{ii0 - 2, jj0 - 2, UW, UH}, bqx{ii0 - 2, jj0, UW, TY}, bqy{ii0, jj0 - 2, TX, UH}, bqe{ii0 - 1, jj0 - 1, TX + 3, TY + 3};
    auto inside = [&](int ii, int jj) { return ii >= 0 && ii < g.NX && jj >= 0 && jj < g.NY; };
    if (ph == 0) {
      CtxBase x; x.g = g;
      for (int c = tid; c < bu.n(); c += NTHR) {
        const int ii = bu.x0 + c % bu.w, jj = bu.y0 + c / bu.w;
        double a = 0.0;
        if (inside(ii, jj)) { x.setpos(ii, jj, kk, tile, i0, j0); a = LDG(aout.v + x.off(aout.nk, 0, 0, 0)); }
        s.oad[c] = a;
      }
    } else if (ph == 1) {
      // reverse of S_a2b_q2: the adjoints of qx, qy and qe where the last phase needs them
      TCtxLinAD<S_a2b_q2> x; x.g = g; x.m = m; x.oad[0] = s.oad; x.ob[0] = bu;
      auto fast = [&](int ii, int jj) {    // (S_a2b_q2::adjoint: uniform weights, qe not read)
        x.setpos(ii, jj, kk, tile, i0, j0);
        return x.i >= 5 && x.i <= g.npx - 5 && x.j >= 5 && x.j <= g.npy - 5 && x.il - 1 >= g.is && x.il + 2 <= g.ie + 1 && x.jl - 1 >= g.js && x.jl + 2 <= g.je + 1;
      };
      for (int c = tid; c < bqx.n(); c += NTHR) {
        const int ii = bqx.x0 + c % bqx.w, jj = bqx.y0 + c / bqx.w;
        double a = 0.0;
        if (inside(ii, jj)) {
          if (fast(ii, jj)) {
            const double* ap = s.oad + bu.idx(ii, jj);
            a = 0.5 * (a2b::a2 * (ap[-UW] + ap[2 * UW]) + a2b::a1 * (ap[0] + ap[UW]));
          } else { x.acc = 0.0; TileLinGather<S_a2b_q2, 0>::run(x, {0}, ii, jj, kk, tile, i0, j0); a = x.acc; }
        }
        s.qx_ad[c] = a;
      }
      for (int c = tid; c < bqy.n(); c += NTHR) {
        const int ii = bqy.x0 + c % bqy.w, jj = bqy.y0 + c / bqy.w;
        double a = 0.0;
        if (inside(ii, jj)) {
          if (fast(ii, jj)) {
            const double* ap = s.oad + bu.idx(ii, jj);
            a = 0.5 * (a2b::a2 * (ap[-1] + ap[2]) + a2b::a1 * (ap[0] + ap[1]));
          } else { x.acc = 0.0; TileLinGather<S_a2b_q2, 1>::run(x, {0}, ii, jj, kk, tile, i0, j0); a = x.acc; }
        }
        s.qy_ad[c] = a;
      }
      for (int c = tid; c < bqe.n(); c += NTHR) {
        const int ii = bqe.x0 + c % bqe.w, jj = bqe.y0 + c / bqe.w;
        double a = 0.0;
        if (inside(ii, jj)) {
          x.setpos(ii, jj, kk, tile, i0, j0);
          if (x.i == 1 || x.i == g.npx || x.j == 1 || x.j == g.npy) {     // qe only exists on the tile boundary
            x.acc = 0.0; TileLinGather<S_a2b_q2, 2>::run(x, {0}, ii, jj, kk, tile, i0, j0); a = x.acc;
          }
        }
        s.qe_ad[c] = a;
      }
    } else {
      auto own = [&](int ii, int jj) {
      if (!inside(ii, jj) || !qin_ad.v) return;
      double acc = 0.0;
      CtxBase xb; xb.g = g; xb.setpos(ii, jj, kk, tile, i0, j0);
      {   // reverse of qx = S_a2b_q1<0>(qin)  (fast path: S_a2b_q1::adjoint)
        const bool fast = xb.i - 1 >= 4 && xb.i + 2 <= g.npx - 3 && xb.il - 1 >= g.is && xb.il + 2 <= g.ie + 1 &&
                          xb.jl >= g.js - 2 && xb.jl <= g.je + 2 && xb.j >= 1 && xb.j <= g.npy - 1;
        if (fast) { const double* ap = s.qx_ad + bqx.idx(ii, jj); acc += a2b::b2 * (ap[-1] + ap[2]) + a2b::b1 * (ap[0] + ap[1]); }
        else {
          TCtxLinAD<S_a2b_q1<0>> x; x.g = g; x.m = m; x.oad[0] = s.qx_ad; x.ob[0] = bqx; x.acc = 0.0;
          TileLinGather<S_a2b_q1<0>, 0>::run(x, {0}, ii, jj, kk, tile, i0, j0); acc += x.acc;
        }
      }
      {
        const bool fast = xb.j - 1 >= 4 && xb.j + 2 <= g.npy - 3 && xb.jl - 1 >= g.js && xb.jl + 2 <= g.je + 1 &&
                          xb.il >= g.is - 2 && xb.il <= g.ie + 2 && xb.i >= 1 && xb.i <= g.npx - 1;
        if (fast) { const double* ap = s.qy_ad + bqy.idx(ii, jj); acc += a2b::b2 * (ap[-TX] + ap[2 * TX]) + a2b::b1 * (ap[0] + ap[TX]); }
        else {
          TCtxLinAD<S_a2b_q1<1>> x; x.g = g; x.m = m; x.oad[0] = s.qy_ad; x.ob[0] = bqy; x.acc = 0.0;
          TileLinGather<S_a2b_q1<1>, 0>::run(x, {0}, ii, jj, kk, tile, i0, j0); acc += x.acc;
        }
      }
      if (!(xb.i > 3 && xb.i < g.npx - 3 && xb.j > 3 && xb.j < g.npy - 3)) {      // (S_a2b_edge::adjoint: boundary outputs read at most two cells inwards)
        TCtxLinAD<S_a2b_edge> x; x.g = g; x.m = m; x.oad[0] = s.qe_ad; x.ob[0] = bqe; x.acc = 0.0;
        TileLinGather<S_a2b_edge, 0>::run(x, {0}, ii, jj, kk, tile, i0, j0); acc += x.acc;
      }
      if (acc != 0.0) { xb.setpos(ii, jj, kk, tile, i0, j0); qin_ad.v[xb.off(qin_ad.nk, 0, 0, 0)] += acc; }
      };
      for (int c = tid; c < TX * TY; c += NTHR) own(ii0 + c % TX, jj0 + c / TX);
    }
  }
};
}  // namespace ftp

// =====================================================================================================================
// a2b_ord4 as a COMPILED LINEAR OPERATOR (FV3LM_FUSED_A2B=2, the default).
//
// a2b_ord4 is linear in qin with purely geometric coefficients (a2b_edge_tlm.F90:546: the tangent is the operator itself, the adjoint
// its transpose).  Three cells away from the cube edges it is the constant 4 x 4 stencil
//     qout(i,j) = sum_ab 0.5 (B_a A_b + A_a B_b) qin(i+a-2, j+b-2),   A = (a2, a1, a1, a2), B = (b2, b1, b1, b2)
// (the two 1-D interpolations of :108-227 composed); near the edges and corners the rows differ cell by cell.  Instead of evaluating the
// edge formulas per thread in every sweep (the stage chain spends 4 launches and ~600 instructions per cell on them everywhere), the
// operator is compiled ONCE per handle: the generic stage chain above is run on 64 probe fields (unit impulses with period 8 in i and j,
// one per "level"), which yields every row of the matrix; rows inside the regular rectangle are checked against the constant stencil,
// the others are kept as sparse rows (and, transposed, as sparse columns for the adjoint).  A sweep is then
//     1 launch over the regular rectangle (16 loads, 3 multiplies per cell) + 1 launch over the sparse rows (one thread per row and level),
// one read and one write of the field in each of NL / TL / AD.  The compiled operator is verified against the chain on a random field
// before it is used (compile_a2b throws otherwise), so a geometry the derivation above does not cover cannot slip through.
// =====================================================================================================================
namespace a2bc {
constexpr int PER = 8, NPROBE = PER * PER;
constexpr double WC = a2b::a2 * a2b::b2, WE = 0.5 * (a2b::a1 * a2b::b2 + a2b::a2 * a2b::b1), WM = a2b::a1 * a2b::b1;

struct Rects { short x0[MAXSUB], x1[MAXSUB], y0[MAXSUB], y1[MAXSUB]; };
struct Tables {
  int nf = 0, nt = 0;
  int *f_ptr = nullptr, *f_tile = nullptr, *f_pos = nullptr, *f_src = nullptr; double* f_w = nullptr;
  int *t_ptr = nullptr, *t_tile = nullptr, *t_pos = nullptr, *t_src = nullptr; double* t_w = nullptr;
  Rects ro, ri;          // regular rectangles of the outputs / of the inputs (transposed operator), array coordinates; empty: x1 < x0
  std::vector<void*> owned;
  ~Tables() { for (void* p : owned) dev::free_(p); }
  template <class T> T* up(const std::vector<T>& v) {
    T* d = (T*)dev::alloc(sizeof(T) * std::max<size_t>(v.size(), 1));
    if (!v.empty()) dev::h2d(d, v.data(), sizeof(T) * v.size());
    owned.push_back(d);
    return d;
  }
};

// q points at the cell; the taps span O0 .. O0+3 in both directions (O0 = -2: the operator, O0 = -1: its transpose)
template <int O0> DEV double stencil(const double* q, int pitch) {
  const double *r0 = q + O0 * pitch + O0, *r1 = r0 + pitch, *r2 = r1 + pitch, *r3 = r2 + pitch;
  const double corners = (r0[0] + r0[3]) + (r3[0] + r3[3]);
  const double edges = ((r0[1] + r0[2]) + (r3[1] + r3[2])) + ((r1[0] + r1[3]) + (r2[0] + r2[3]));
  const double centers = (r1[1] + r1[2]) + (r2[1] + r2[2]);
  return WC * corners + WE * edges + WM * centers;
}
template <bool TL> struct KReg {      // regular outputs
  Rects r; int pitch, slab, nk; const double* in; const double* ind; double* out; double* outd;
  DEV void operator()(int ii, int jj, int z) const {
    int tile, kk; split_z(z, nk, tile, kk);
    if (ii < r.x0[tile] || ii > r.x1[tile] || jj < r.y0[tile] || jj > r.y1[tile]) return;
    const int p = (tile * nk + kk) * slab + jj * pitch + ii;
    out[p] = stencil<-2>(in + p, pitch);
    if (TL) outd[p] = stencil<-2>(ind + p, pitch);
  }
};
struct KRegT {                        // regular inputs of the transposed operator
  Rects r; int pitch, slab, nk; const double* oad; double* iad;
  DEV void operator()(int ii, int jj, int z) const {
    int tile, kk; split_z(z, nk, tile, kk);
    if (ii < r.x0[tile] || ii > r.x1[tile] || jj < r.y0[tile] || jj > r.y1[tile]) return;
    const int p = (tile * nk + kk) * slab + jj * pitch + ii;
    iad[p] += stencil<-1>(oad + p, pitch);
  }
};
template <bool TL> struct KRows {     // sparse rows: one thread per (row, level)
  const int *ptr, *tile, *pos, *src; const double* w; int slab, nk; const double* in; const double* ind; double* out; double* outd;
  DEV void operator()(int r, int k, int) const {
    const int base = (tile[r] * nk + k) * slab;
    double s = 0.0, sd = 0.0;
    for (int q = ptr[r]; q < ptr[r + 1]; q++) { s += w[q] * in[base + src[q]]; if (TL) sd += w[q] * ind[base + src[q]]; }
    out[base + pos[r]] = s;
    if (TL) outd[base + pos[r]] = sd;
  }
};
struct KRowsT {
  const int *ptr, *tile, *pos, *src; const double* w; int slab, nk; const double* oad; double* iad;
  DEV void operator()(int r, int k, int) const {
    const int base = (tile[r] * nk + k) * slab;
    double s = 0.0;
    for (int q = ptr[r]; q < ptr[r + 1]; q++) s += w[q] * oad[base + src[q]];
    iad[base + pos[r]] += s;
  }
};

static void add_chain(Program& T, int qin, int qout, int nk) {
  int qx = T.val("qx", nk), qy = T.val("qy", nk), qe = T.val("qe", nk);
  T.add<S_a2b_q1<0>>("a2b_qx", {0}, {qin}, {qx}, nk);
  T.add<S_a2b_q1<1>>("a2b_qy", {0}, {qin}, {qy}, nk);
  T.add<S_a2b_edge>("a2b_edge", {0}, {qin}, {qe}, nk);
  T.add<S_a2b_q2>("a2b_q2", {0}, {qx, qy, qe}, {qout}, nk);
}

static std::shared_ptr<Tables> compile(Device* dv) {
  const Geom& g = dv->g;
  const int NP = NPROBE + 1, lo = g.ng - 1;          // the last "level" is a random field: the check of the compiled operator
  Program T; T.dv = dv; T.name = "a2b_probe";
  int qin = T.val("qin", NP, true), qout = T.val("qout", NP, true);
  add_chain(T, qin, qout, NP);
  const size_t n = T.val_doubles(qin);
  std::vector<double> hin(n, 0.0), hout(n, 0.0);
  auto at = [&](int t, int p, int jj, int ii) { return (((size_t)t * NP + p) * g.NY + jj) * g.pitch + ii; };
  unsigned long long seed = 88172645463325252ULL;
  auto rnd = [&]() { seed ^= seed << 13; seed ^= seed >> 7; seed ^= seed << 17; return (double)(seed >> 11) / 9007199254740992.0 - 0.5; };
  for (int t = 0; t < g.ntile; t++)
    for (int jj = 0; jj < g.NY; jj++)
      for (int ii = 0; ii < g.NX; ii++) {
        hin[at(t, (jj % PER) * PER + ii % PER, jj, ii)] = 1.0;
        hin[at(t, NPROBE, jj, ii)] = rnd();
      }
  double* din = dv->pool.get(n); double* dout = dv->pool.get(n);
  dev::h2d(din, hin.data(), n * sizeof(double));
  dev::zero(dout, n * sizeof(double));
  T.vals[qin].traj = din; T.vals[qout].traj = dout;
  try { T.run(MODE_NL); dev::sync(); dev::d2h(hout.data(), dout, n * sizeof(double)); }
  catch (...) { dv->pool.put(din); dv->pool.put(dout); throw; }
  dv->pool.put(din); dv->pool.put(dout);

  auto tb = std::make_shared<Tables>();
  std::vector<int> f_ptr{0}, f_tile, f_pos, f_src, t_ptr{0}, t_tile, t_pos, t_src;
  std::vector<double> f_w, t_w;
  struct Ent { int cell; double w; };
  const double tol = 1e-13;
  for (int t = 0; t < g.ntile; t++) {
    const int ci = g.i0[t] - lo, cj = g.j0[t] - lo;
    const int xs = g.is + lo, xe = g.ie + lo + 1, ys = g.js + lo, ye = g.je + lo + 1;      // computed outputs (is..ie+1, js..je+1)
    // every row of the operator from the impulse responses
    std::vector<std::vector<Ent>> rows((size_t)g.NY * g.NX), cols((size_t)g.NY * g.NX);
    double worst = 0.0;
    for (int jj = ys; jj <= ye; jj++)
      for (int ii = xs; ii <= xe; ii++) {
        auto& row = rows[(size_t)jj * g.NX + ii];
        double chk = 0.0;
        for (int p = 0; p < NPROBE; p++) {
          const double v = hout[at(t, p, jj, ii)];
          if (v == 0.0) continue;
          const int a = p % PER, b = p / PER;
          const int si = ii - 4 + (((a - (ii - 4)) % PER) + PER) % PER, sj = jj - 4 + (((b - (jj - 4)) % PER) + PER) % PER;
          if (si < 0 || si >= g.NX || sj < 0 || sj >= g.NY) throw std::runtime_error("a2b_ord4 compile: an impulse response outside the array");
          row.push_back({sj * g.NX + si, v});
          chk += v * hin[at(t, NPROBE, sj, si)];
        }
        const double ref = hout[at(t, NPROBE, jj, ii)];
        worst = std::max(worst, fabs(chk - ref) / std::max(1.0, fabs(ref)));
      }
    if (worst > 1e-12) throw std::runtime_error("a2b_ord4 compile: the probed operator does not reproduce the stage chain (footprint wider than the probe period?)");
    // regular outputs: faces 3 .. npx-2 / npy-2 of the cube tile; every row there must be the constant stencil
    int ox0 = std::max(xs, 3 - ci), ox1 = std::min(xe, g.npx - 2 - ci), oy0 = std::max(ys, 3 - cj), oy1 = std::min(ye, g.npy - 2 - cj);
    auto weight = [](int d0, int d1) {   // d = 0..3 along each direction
      const bool e0 = d0 == 0 || d0 == 3, e1 = d1 == 0 || d1 == 3;
      return e0 && e1 ? WC : (e0 || e1 ? WE : WM);
    };
    auto is_stencil = [&](const std::vector<Ent>& r, int ii, int jj, int o0) {
      if (r.size() != 16) return false;
      for (const Ent& e : r) {
        const int di = e.cell % g.NX - ii - o0, dj = e.cell / g.NX - jj - o0;
        if (di < 0 || di > 3 || dj < 0 || dj > 3 || fabs(e.w - weight(di, dj)) > tol) return false;
      }
      return true;
    };
    for (int jj = oy0; jj <= oy1; jj++)
      for (int ii = ox0; ii <= ox1; ii++)
        if (!is_stencil(rows[(size_t)jj * g.NX + ii], ii, jj, -2)) throw std::runtime_error("a2b_ord4 compile: a row of the regular rectangle is not the constant stencil");
    tb->ro.x0[t] = (short)ox0; tb->ro.x1[t] = (short)ox1; tb->ro.y0[t] = (short)oy0; tb->ro.y1[t] = (short)oy1;
    for (int jj = ys; jj <= ye; jj++)
      for (int ii = xs; ii <= xe; ii++) {
        const auto& row = rows[(size_t)jj * g.NX + ii];
        for (const Ent& e : row) cols[e.cell].push_back({jj * g.NX + ii, e.w});
        if (ii >= ox0 && ii <= ox1 && jj >= oy0 && jj <= oy1) continue;
        f_tile.push_back(t); f_pos.push_back(jj * g.pitch + ii);
        for (const Ent& e : row) { f_src.push_back((e.cell / g.NX) * g.pitch + e.cell % g.NX); f_w.push_back(e.w); }
        f_ptr.push_back((int)f_src.size());
      }
    // regular inputs of the transpose: all sixteen readers are regular outputs and nothing else reads the cell
    int ix0 = ox0 + 1, ix1 = ox1 - 2, iy0 = oy0 + 1, iy1 = oy1 - 2;
    for (int shrink = 0; shrink < 8; shrink++) {
      bool ok = true;
      for (int jj = iy0; jj <= iy1 && ok; jj++)
        for (int ii = ix0; ii <= ix1 && ok; ii++) ok = is_stencil(cols[(size_t)jj * g.NX + ii], ii, jj, -1);
      if (ok) break;
      ix0++; ix1--; iy0++; iy1--;
      if (shrink == 7) { ix1 = ix0 - 1; iy1 = iy0 - 1; }
    }
    tb->ri.x0[t] = (short)ix0; tb->ri.x1[t] = (short)ix1; tb->ri.y0[t] = (short)iy0; tb->ri.y1[t] = (short)iy1;
    for (int jj = 0; jj < g.NY; jj++)
      for (int ii = 0; ii < g.NX; ii++) {
        const auto& col = cols[(size_t)jj * g.NX + ii];
        if (col.empty() || (ii >= ix0 && ii <= ix1 && jj >= iy0 && jj <= iy1)) continue;
        t_tile.push_back(t); t_pos.push_back(jj * g.pitch + ii);
        for (const Ent& e : col) { t_src.push_back((e.cell / g.NX) * g.pitch + e.cell % g.NX); t_w.push_back(e.w); }
        t_ptr.push_back((int)t_src.size());
      }
  }
  for (int t = g.ntile; t < MAXSUB; t++) { tb->ro.x0[t] = tb->ri.x0[t] = 1; tb->ro.x1[t] = tb->ri.x1[t] = 0; tb->ro.y0[t] = tb->ri.y0[t] = 1; tb->ro.y1[t] = tb->ri.y1[t] = 0; }
  tb->nf = (int)f_tile.size(); tb->nt = (int)t_tile.size();
  tb->f_ptr = tb->up(f_ptr); tb->f_tile = tb->up(f_tile); tb->f_pos = tb->up(f_pos); tb->f_src = tb->up(f_src); tb->f_w = tb->up(f_w);
  tb->t_ptr = tb->up(t_ptr); tb->t_tile = tb->up(t_tile); tb->t_pos = tb->up(t_pos); tb->t_src = tb->up(t_src); tb->t_w = tb->up(t_w);
  return tb;
}
}  // namespace a2bc

int build_a2b_ord4(Program& P, Mosaic& mo, int qin, int nk, const std::string& tag) {
  (void)mo;
  auto nm = [&](const char* s) { return tag + "." + s; };
  const char* fe = getenv("FV3LM_FUSED_A2B");
  const int level = fe ? atoi(fe) : 2;
  if (level == 2 && P.name != "a2b_probe") {
    if (P.vals[qin].nk != nk) throw std::runtime_error("compiled a2b_ord4: the field must have the launch's number of levels");
    int qout = P.val(nm("qout"), nk);
    Op op; op.name = "a2b_compiled"; op.in = {qin}; op.out = {qout}; op.nk_launch = nk; op.tl_only = P.tl_only; op.variant = P.variant;
    op.run = [](Program& P, Op& o, int mode) {
      // compiled on first use (the metrics must be on the device; the first run of a step program is eager, so the probe sweep and its
      // host round trip never fall into a graph capture)
      if (!P.dv->a2b_tables) P.dv->a2b_tables = a2bc::compile(P.dv);
      const a2bc::Tables* tb = static_cast<const a2bc::Tables*>(P.dv->a2b_tables.get());
      const Geom& g = P.dv->g;
      const Value &vi = P.vals[o.in[0]], &vo = P.vals[o.out[0]];
      const int nk = o.nk_launch;
      if (mode == MODE_AD) {
        if (!vi.active || !vo.active || !vo.pert || !vi.pert) return;
        launch3d(a2bc::KRegT{tb->ri, g.pitch, g.slab, nk, vo.pert, vi.pert}, g.NX, g.NY, g.ntile * nk);
        if (tb->nt) launch3d(a2bc::KRowsT{tb->t_ptr, tb->t_tile, tb->t_pos, tb->t_src, tb->t_w, g.slab, nk, vo.pert, vi.pert}, tb->nt, nk, 1);
      } else if (mode == MODE_TL && vi.active && vi.pert && vo.pert) {
        launch3d(a2bc::KReg<true>{tb->ro, g.pitch, g.slab, nk, vi.traj, vi.pert, vo.traj, vo.pert}, g.NX, g.NY, g.ntile * nk);
        if (tb->nf) launch3d(a2bc::KRows<true>{tb->f_ptr, tb->f_tile, tb->f_pos, tb->f_src, tb->f_w, g.slab, nk, vi.traj, vi.pert, vo.traj, vo.pert}, tb->nf, nk, 1);
      } else {
        launch3d(a2bc::KReg<false>{tb->ro, g.pitch, g.slab, nk, vi.traj, nullptr, vo.traj, nullptr}, g.NX, g.NY, g.ntile * nk);
        if (tb->nf) launch3d(a2bc::KRows<false>{tb->f_ptr, tb->f_tile, tb->f_pos, tb->f_src, tb->f_w, g.slab, nk, vi.traj, nullptr, vo.traj, nullptr}, tb->nf, nk, 1);
      }
    };
    P.ops.push_back(op);
    return qout;
  }
  if (level == 1) {
    if (P.vals[qin].nk != nk) throw std::runtime_error("fused a2b_ord4: the field must have the launch's number of levels");
    int qout = P.val(nm("qout"), nk);
    Op op; op.name = "a2b_fused"; op.in = {qin}; op.out = {qout}; op.nk_launch = nk; op.tl_only = P.tl_only; op.variant = P.variant;
    op.run = [](Program& P, Op& o, int mode) {
      const Geom& g = P.dv->g;
      const Value &vi = P.vals[o.in[0]], &vo = P.vals[o.out[0]];
      if (mode == MODE_AD) {
        if (!vi.active || !vo.active || !vo.pert) return;
        ftp::KernA2bRev k{}; k.g = g; k.m = P.dv->m; k.nk = o.nk_launch;
        k.aout = ftp::adj_in(vo); k.qin_ad = ftp::adj_out(vi);
        ftp::launch_tile(k, g.NX, g.NY, g.ntile * o.nk_launch);
      } else if (mode == MODE_TL && vi.active && vi.pert) {
        ftp::KernA2b<Dual> k{}; k.g = g; k.m = P.dv->m; k.nk = o.nk_launch; k.qin = ftp::fld(vi, true); k.qout = ftp::ofld(vo, true);
        ftp::launch_tile(k, g.NX, g.NY, g.ntile * o.nk_launch);
      } else {
        ftp::KernA2b<double> k{}; k.g = g; k.m = P.dv->m; k.nk = o.nk_launch; k.qin = ftp::fld(vi, false); k.qout = ftp::ofld(vo, false);
        ftp::launch_tile(k, g.NX, g.NY, g.ntile * o.nk_launch);
      }
    };
    P.ops.push_back(op);
    return qout;
  }
  int qx = P.val(nm("qx"), nk), qy = P.val(nm("qy"), nk), qe = P.val(nm("qe"), nk), qout = P.val(nm("qout"), nk);
  P.add<S_a2b_q1<0>>("a2b_qx", {0}, {qin}, {qx}, nk);
  P.add<S_a2b_q1<1>>("a2b_qy", {0}, {qin}, {qy}, nk);
  P.add<S_a2b_edge>("a2b_edge", {0}, {qin}, {qe}, nk);
  P.add<S_a2b_q2>("a2b_q2", {0}, {qx, qy, qe}, {qout}, nk);
  return qout;
}

// host: the 12 extrap_corner weights of every resident sub-domain that holds a tile corner
// (lon/lat arrays are the local [ntile][NY][NX] host copies of grid and agrid)
void a2b_corner_weights(const Geom& g, const double* glon, const double* glat, const double* alon, const double* alat, double* out /* [ntile][12] */) {
  const int o = g.ng - 1, npx = g.npx, npy = g.npy;
  auto gc = [](double lon1, double lat1, double lon2, double lat2) {
    double s1 = sin(0.5 * (lat1 - lat2)), s2 = sin(0.5 * (lon1 - lon2));
    return 2.0 * asin(sqrt(s1 * s1 + cos(lat1) * cos(lat2) * s2 * s2));
  };
  // {corner point, a1, a2} in tile-global indices, same order as S_a2b_edge::eval
  const int T[12][6] = {{1, 1, 1, 1, 2, 2}, {1, 1, 0, 1, -1, 2}, {1, 1, 1, 0, 2, -1},
                        {npx, 1, npx - 1, 1, npx - 2, 2}, {npx, 1, npx - 1, 0, npx - 2, -1}, {npx, 1, npx, 1, npx + 1, 2},
                        {npx, npy, npx - 1, npy - 1, npx - 2, npy - 2}, {npx, npy, npx, npy - 1, npx + 1, npy - 2}, {npx, npy, npx - 1, npy, npx - 2, npy + 1},
                        {1, npy, 1, npy - 1, 2, npy - 2}, {1, npy, 0, npy - 1, -1, npy - 2}, {1, npy, 1, npy, 2, npy + 1}};
  for (int t = 0; t < g.ntile; t++)
    for (int n = 0; n < 12; n++) {
      out[t * 12 + n] = 0.0;
      auto at = [&](const double* a, int i, int j, bool& ok) {
        int ii = i - g.i0[t] + o, jj = j - g.j0[t] + o;
        if (ii < 0 || ii >= g.NX || jj < 0 || jj >= g.NY) { ok = false; return 0.0; }
        return a[((size_t)t * g.NY + jj) * g.NX + ii];
      };
      bool ok = true;
      // the corner point must be one this sub-domain computes (is..ie+1, js..je+1)
      int il = T[n][0] - g.i0[t], jl = T[n][1] - g.j0[t];
      if (il < 1 || il > g.ie + 1 || jl < 1 || jl > g.je + 1) continue;
      double l0 = at(glon, T[n][0], T[n][1], ok), t0 = at(glat, T[n][0], T[n][1], ok);
      double x1 = gc(at(alon, T[n][2], T[n][3], ok), at(alat, T[n][2], T[n][3], ok), l0, t0);
      double x2 = gc(at(alon, T[n][4], T[n][5], ok), at(alat, T[n][4], T[n][5], ok), l0, t0);
      if (ok) out[t * 12 + n] = x1 / (x2 - x1);
    }
}

void mod_a2b_ord4(Program& P, Mosaic& mo, ModuleIO& io, const ModuleParams&) {
  const int K = P.dv->g.K;
  int qin = io.in(P, "qin", K);
  io.out(P, "qout", build_a2b_ord4(P, mo, qin, K, "a2b"));
}

}  // namespace fv3lm
