// a2b_ord4: 4th-order A-grid -> B-grid interpolation (model/a2b_edge_nlm.F90:49-329,
// extrap_corner :800-810; TL a2b_edge_tlm.F90:546, AD a2b_edge_adm.F90:72).  Linear in qin.
#include "stages_dsw.h"
#include "modules.h"

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

int build_a2b_ord4(Program& P, Mosaic& mo, int qin, int nk, const std::string& tag) {
  (void)mo;
  auto nm = [&](const char* s) { return tag + "." + s; };
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
