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

int build_a2b_ord4(Program& P, Mosaic& mo, int qin, int nk, const std::string& tag) {
  (void)mo;
  auto nm = [&](const char* s) { return tag + "." + s; };
  const char* fe = getenv("FV3LM_FUSED_A2B");
  if (fe && atoi(fe) != 0) {
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
