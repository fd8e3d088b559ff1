// dyn_core helpers: geopk (model/dyn_core_nlm.F90:1954-2087), p_grad_c (:1369-1429),
// one_grad_p (:1645-1779), nh_p_grad (:1431-1529) and small accumulation stages.
// TL: model_tlmadm/dyn_core_tlm.F90 GEOPK_TLM :4578, P_GRAD_C_TLM :3194, ONE_GRAD_P_TLM :3867,
// NH_P_GRAD_TLM :3340; AD: dyn_core_adm.F90.
#pragma once
#include "engine.h"
#include "mosaic.h"

namespace fv3lm {

// geopk: column prefix sum of delp -> pe, peln, pk = exp(akap*log(pe)); suffix sum of
// cp*pt*dpk -> gz; pkz.   in: delp pt hs ; out: pk gz pe peln pkz
struct S_geopk {
  static constexpr int NI = 3, NO = 5;
  struct P { double ptop, akap, cp_air; int halo; int cg; int K; };   // halo = 1 (C grid) or 2 (D grid)
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is - p.halo, g.ie + p.halo, g.js - p.halo, g.je + p.halo)) return;
    const int K = p.K;
    const double ptk = pow(p.ptop, p.akap), peln1 = log(p.ptop);
    const bool in_dom = x.in_rect(g.is, g.ie, g.js, g.je);
    T p1 = T(p.ptop);
    x.out(0, 0, T(ptk)); x.out(2, 0, T(p.ptop)); x.out(3, 0, T(peln1));
    T pk_prev = T(ptk), ln_prev = T(peln1);
    for (int k = 1; k <= K; k++) {
      p1 = p1 + x.in(0, k - 1);
      T lp = m_log(p1);
      T pk = m_exp(p.akap * lp);
      x.out(0, k, pk); x.out(2, k, p1); x.out(3, k, lp);
      if (in_dom && !p.cg) x.out(4, k - 1, (pk - pk_prev) / (p.akap * (lp - ln_prev)));
      pk_prev = pk; ln_prev = lp;
    }
    T gz = x.in(2, 0);
    x.out(1, K, gz);
    for (int k = K - 1; k >= 0; k--) {
      gz = gz + p.cp_air * x.in(1, k) * (x.rd(0, k + 1) - x.rd(0, k));
      x.out(1, k, gz);
    }
  }
  template <class X> DEV static void eval_ad(X& x, const P& p) {
    const Geom& g = x.g;
    if (!x.in_rect(g.is - p.halo, g.ie + p.halo, g.js - p.halo, g.je + p.halo)) return;
    const int K = p.K;
    const bool in_dom = x.in_rect(g.is, g.ie, g.js, g.je);
    // pkz -> pk, peln adjoints (the output adjoint arrays serve as workspace; they are dead afterwards)
    if (in_dom && !p.cg) {
      for (int k = 0; k < K; k++) {
        double a = x.oad(4, k);
        double dln = x.outv(3, k + 1) - x.outv(3, k), pkz = x.outv(4, k);
        double dpk_ad = a / (p.akap * dln), dln_ad = -a * pkz / dln;
        x.oad_add(0, k + 1, dpk_ad); x.oad_add(0, k, -dpk_ad);
        x.oad_add(3, k + 1, dln_ad); x.oad_add(3, k, -dln_ad);
      }
    }
    // gz suffix sum:  gz(k) = gz(k+1) + cp*pt(k)*(pk(k+1)-pk(k))
    double g_ad = 0.0;
    for (int k = 0; k < K; k++) {
      g_ad += x.oad(1, k);
      double dpk = x.outv(0, k + 1) - x.outv(0, k);
      x.add(1, k, g_ad * p.cp_air * dpk);
      double c = g_ad * p.cp_air * x.in(1, k);
      x.oad_add(0, k + 1, c); x.oad_add(0, k, -c);
    }
    // pressure prefix sum:  p(k) = p(k-1) + delp(k-1); lp = log p; pk = exp(akap lp)
    double p_ad = 0.0;
    for (int k = K; k >= 1; k--) {
      double lp_ad = x.oad(0, k) * p.akap * x.outv(0, k) + x.oad(3, k);
      p_ad += lp_ad / x.outv(2, k) + x.oad(2, k);
      x.add(0, k - 1, p_ad);
    }
  }
};

// p_grad_c (hydrostatic: wk = pkc(k+1)-pkc(k); non-hydrostatic: wk = delpc)
// in: uc vc pkc gz delpc ; out: uc_new vc_new
struct S_pgrad_c {
  static constexpr int NI = 5, NO = 2;
  struct P { double dt2; int hydrostatic; };
  static constexpr int NT = 17;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0},
                                   {2, 0, 0, 0}, {2, -1, 0, 0}, {2, 0, -1, 0}, {2, 0, 0, 1}, {2, -1, 0, 1}, {2, 0, -1, 1},
                                   {3, 0, 0, 0}, {3, -1, 0, 0}, {3, 0, -1, 0}, {3, 0, 0, 1}, {3, -1, 0, 1}, {3, 0, -1, 1},
                                   {4, 0, 0, 0}, {4, -1, 0, 0}, {4, 0, -1, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    auto wk = [&](int di, int dj) -> T { return p.hydrostatic ? x.in(2, di, dj, 1) - x.in(2, di, dj, 0) : x.in(4, di, dj); };
    if (x.in_rect(g.is, g.ie + 1, g.js, g.je)) {
      x.out(0, x.in(0) + p.dt2 * x.M(x.m.rdxc) / (wk(-1, 0) + wk(0, 0)) *
                             ((x.in(3, -1, 0, 1) - x.in(3, 0, 0, 0)) * (x.in(2, 0, 0, 1) - x.in(2, -1, 0, 0)) +
                              (x.in(3, -1, 0, 0) - x.in(3, 0, 0, 1)) * (x.in(2, -1, 0, 1) - x.in(2, 0, 0, 0))));
    }
    if (x.in_rect(g.is, g.ie, g.js, g.je + 1)) {
      x.out(1, x.in(1) + p.dt2 * x.M(x.m.rdyc) / (wk(0, -1) + wk(0, 0)) *
                             ((x.in(3, 0, -1, 1) - x.in(3, 0, 0, 0)) * (x.in(2, 0, 0, 1) - x.in(2, 0, -1, 0)) +
                              (x.in(3, 0, -1, 0) - x.in(3, 0, 0, 1)) * (x.in(2, 0, -1, 1) - x.in(2, 0, 0, 0))));
    }
  }

  // ---- hand-derived gather adjoint, same structure as S_gradp::adjoint below: the output at (o, k) reads the corners
  //   a = (o - e, k)   b = (o, k)   c = (o - e, k+1)   d = (o, k+1),   e = (1,0) for uc, (0,1) for vc
  //   out = wind + dt2 rdc (A B + C D) / den,  A = gz_c - gz_b, B = pk_d - pk_a, C = gz_a - gz_d, D = pk_c - pk_b,
  //   den = (pk_c - pk_a) + (pk_d - pk_b) (hydrostatic)  or  delpc_a + delpc_b
  static constexpr bool custom_ad = true;
  template <class K> DEV static void adjoint(const K& kn, int ii, int jj, int kk, int tile, double* acc) {
    const P& p = kn.p;
    const Geom& g = kn.g;
    CtxNL<S_pgrad_c> x; x.g = kn.g; x.m = kn.m; x.in_ = kn.in;
    const int i0 = g.i0[tile], j0 = g.j0[tile];
    if (kk < kn.nk_fwd) {
      x.setpos(ii, jj, kk, tile, i0, j0);
      if (kn.outad.p[0] && x.in_rect(g.is, g.ie + 1, g.js, g.je)) acc[0] += kn.outad.p[0][x.off(kn.outad.nk[0], 0, 0, 0)];
      if (kn.outad.p[1] && x.in_rect(g.is, g.ie, g.js, g.je + 1)) acc[1] += kn.outad.p[1][x.off(kn.outad.nk[1], 0, 0, 0)];
    }
    // role: 0 = a, 1 = b, 2 = c, 3 = d of the output at (oi, oj, ok)
    auto side = [&](const bool isu, const int role) {
      const int oi = ii + ((isu && !(role & 1)) ? 1 : 0), oj = jj + ((!isu && !(role & 1)) ? 1 : 0), ok = kk - (role >> 1);
      if (oi < 0 || oj < 0 || oi >= g.NX || oj >= g.NY || ok < 0 || ok >= kn.nk_fwd) return;
      const int o = isu ? 0 : 1;
      if (!kn.outad.p[o]) return;
      x.setpos(oi, oj, ok, tile, i0, j0);
      if (isu ? !x.in_rect(g.is, g.ie + 1, g.js, g.je) : !x.in_rect(g.is, g.ie, g.js, g.je + 1)) return;
      const double au = kn.outad.p[o][x.off(kn.outad.nk[o], 0, 0, 0)];
      if (au == 0.0) return;
      const int ex = isu ? 1 : 0, ey = isu ? 0 : 1;
      const double pa = x.in(2, -ex, -ey, 0), pb = x.in(2, 0, 0, 0), pc = x.in(2, -ex, -ey, 1), pd = x.in(2, 0, 0, 1);
      const double ga = x.in(3, -ex, -ey, 0), gb = x.in(3, 0, 0, 0), gc = x.in(3, -ex, -ey, 1), gd = x.in(3, 0, 0, 1);
      const double A = gc - gb, B = pd - pa, C = ga - gd, D = pc - pb;
      const double den = p.hydrostatic ? (pc - pa) + (pd - pb) : x.in(4, -ex, -ey) + x.in(4, 0, 0);
      const double r = p.dt2 * (isu ? x.M(x.m.rdxc) : x.M(x.m.rdyc)) * au / den, q = (A * B + C * D) / den;
      const double qd = p.hydrostatic ? q : 0.0;         // den depends on pk only in the hydrostatic form
      acc[2] += r * (role == 0 ? (-A + qd) : role == 1 ? (-C + qd) : role == 2 ? (C - qd) : (A - qd));
      acc[3] += r * (role == 0 ? D : role == 1 ? -B : role == 2 ? B : -D);
      if (!p.hydrostatic && role < 2) acc[4] += -r * q;
    };
#pragma unroll
    for (int role = 0; role < 4; role++) { side(true, role); side(false, role); }
  }
};

// one_grad_p / nh_p_grad wind update from B-grid (corner) pk, gz [, pp, delp_b]
// in: u v pkb gzb ppb dpb ; out: u_new v_new.    pk(k=0) = top value (ptk) and pp(k=0) = 0 are
// imposed on read, like the reference does on the arrays (:1689-1693, :1465-1470).
struct S_gradp {
  static constexpr int NI = 6, NO = 2;
  struct P { double dt, top; int nonhydro; };
  static constexpr int NT = 23;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0},
                                   {2, 0, 0, 0}, {2, 1, 0, 0}, {2, 0, 1, 0}, {2, 0, 0, 1}, {2, 1, 0, 1}, {2, 0, 1, 1},
                                   {3, 0, 0, 0}, {3, 1, 0, 0}, {3, 0, 1, 0}, {3, 0, 0, 1}, {3, 1, 0, 1}, {3, 0, 1, 1},
                                   {4, 0, 0, 0}, {4, 1, 0, 0}, {4, 0, 1, 0}, {4, 0, 0, 1}, {4, 1, 0, 1}, {4, 0, 1, 1},
                                   {5, 0, 0, 0}, {5, 1, 0, 0}, {5, 0, 1, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    auto PK = [&](int di, int dj, int dk) -> T { return (x.kk + dk == 0) ? T(p.top) : x.in(2, di, dj, dk); };
    auto PP = [&](int di, int dj, int dk) -> T { return (x.kk + dk == 0) ? T(0.0) : x.in(4, di, dj, dk); };
    auto GZ = [&](int di, int dj, int dk) -> T { return x.in(3, di, dj, dk); };
    auto wk = [&](int di, int dj) -> T { return PK(di, dj, 1) - PK(di, dj, 0); };
    if (x.in_rect(g.is, g.ie, g.js, g.je + 1)) {
      T du = p.dt / (wk(0, 0) + wk(1, 0)) * ((GZ(0, 0, 1) - GZ(1, 0, 0)) * (PK(1, 0, 1) - PK(0, 0, 0)) + (GZ(0, 0, 0) - GZ(1, 0, 1)) * (PK(0, 0, 1) - PK(1, 0, 0)));
      if (p.nonhydro) {
        T dn = p.dt / (x.in(5, 0, 0) + x.in(5, 1, 0)) * ((GZ(0, 0, 1) - GZ(1, 0, 0)) * (PP(1, 0, 1) - PP(0, 0, 0)) + (GZ(0, 0, 0) - GZ(1, 0, 1)) * (PP(0, 0, 1) - PP(1, 0, 0)));
        x.out(0, (x.in(0) + du + dn) * x.M(x.m.rdx));
      } else {
        x.out(0, x.M(x.m.rdx) * (0.0 + x.in(0) + du));
      }
    }
    if (x.in_rect(g.is, g.ie + 1, g.js, g.je)) {
      T dv = p.dt / (wk(0, 0) + wk(0, 1)) * ((GZ(0, 0, 1) - GZ(0, 1, 0)) * (PK(0, 1, 1) - PK(0, 0, 0)) + (GZ(0, 0, 0) - GZ(0, 1, 1)) * (PK(0, 0, 1) - PK(0, 1, 0)));
      if (p.nonhydro) {
        T dn = p.dt / (x.in(5, 0, 0) + x.in(5, 0, 1)) * ((GZ(0, 0, 1) - GZ(0, 1, 0)) * (PP(0, 1, 1) - PP(0, 0, 0)) + (GZ(0, 0, 0) - GZ(0, 1, 1)) * (PP(0, 0, 1) - PP(0, 1, 0)));
        x.out(1, (x.in(1) + dv + dn) * x.M(x.m.rdy));
      } else {
        x.out(1, x.M(x.m.rdy) * (0.0 + x.in(1) + dv));
      }
    }
  }

  // ---- hand-derived gather adjoint (replaces six multi-seed dual evaluations of the whole stage per cell).  A wind output at
  // (o, k) reads the B-grid fields at the four corners of its (horizontal edge) x (layer) face:
  //   a = (o, k)   b = (o + e, k)   c = (o, k+1)   d = (o + e, k+1),   e = (1,0) for u, (0,1) for v
  //   du = dt (A B + C D) / den,  A = gz_c - gz_b, B = pk_d - pk_a, C = gz_a - gz_d, D = pk_c - pk_b, den = (pk_c - pk_a) + (pk_d - pk_b)
  //   dn = dt (A (pp_d - pp_a) + C (pp_c - pp_b)) / (dp_a + dp_b)                                           (non-hydrostatic)
  // so a corner (i, j, k) collects from the u outputs in which it is a, b, c or d and from the four v outputs likewise.
  static constexpr bool custom_ad = true;
  template <class K> DEV static void adjoint(const K& kn, int ii, int jj, int kk, int tile, double* acc) {
    const P& p = kn.p;
    const Geom& g = kn.g;
    CtxNL<S_gradp> x; x.g = kn.g; x.m = kn.m; x.in_ = kn.in;
    const int i0 = g.i0[tile], j0 = g.j0[tile];
    // the winds themselves
    if (kk < kn.nk_fwd) {
      x.setpos(ii, jj, kk, tile, i0, j0);
      if (kn.outad.p[0] && x.in_rect(g.is, g.ie, g.js, g.je + 1)) acc[0] += x.M(x.m.rdx) * kn.outad.p[0][x.off(kn.outad.nk[0], 0, 0, 0)];
      if (kn.outad.p[1] && x.in_rect(g.is, g.ie + 1, g.js, g.je)) acc[1] += x.M(x.m.rdy) * kn.outad.p[1][x.off(kn.outad.nk[1], 0, 0, 0)];
    }
    // role: 0 = a, 1 = b, 2 = c, 3 = d of the output at (oi, oj, ok); isu: a u output (e = (1,0)) or a v output (e = (0,1))
    auto side = [&](const bool isu, const int role) {
      const int oi = ii - ((isu && (role & 1)) ? 1 : 0), oj = jj - ((!isu && (role & 1)) ? 1 : 0), ok = kk - (role >> 1);
      if (oi < 0 || oj < 0 || oi >= g.NX || oj >= g.NY || ok < 0 || ok >= kn.nk_fwd) return;
      const int o = isu ? 0 : 1;
      if (!kn.outad.p[o]) return;
      x.setpos(oi, oj, ok, tile, i0, j0);
      if (isu ? !x.in_rect(g.is, g.ie, g.js, g.je + 1) : !x.in_rect(g.is, g.ie + 1, g.js, g.je)) return;
      const double au = kn.outad.p[o][x.off(kn.outad.nk[o], 0, 0, 0)];
      if (au == 0.0) return;
      const int ex = isu ? 1 : 0, ey = isu ? 0 : 1;
      const bool top = ok == 0;
      const double pa = top ? p.top : x.in(2, 0, 0, 0), pb = top ? p.top : x.in(2, ex, ey, 0), pc = x.in(2, 0, 0, 1), pd = x.in(2, ex, ey, 1);
      const double ga = x.in(3, 0, 0, 0), gb = x.in(3, ex, ey, 0), gc = x.in(3, 0, 0, 1), gd = x.in(3, ex, ey, 1);
      const double A = gc - gb, B = pd - pa, C = ga - gd, D = pc - pb, den = (pc - pa) + (pd - pb);
      const double w = au * (isu ? x.M(x.m.rdx) : x.M(x.m.rdy));
      const double r = p.dt * w / den, q = (A * B + C * D) / den;
      // pk: d(num/den)/dx = (dnum/dx - q dden/dx) / den ;  gz: dnum/dx / den
      const double dpk = role == 0 ? (-A + q) : role == 1 ? (-C + q) : role == 2 ? (C - q) : (A - q);
      const double dgz = role == 0 ? D : role == 1 ? -B : role == 2 ? B : -D;
      if (!(top && role < 2)) acc[2] += r * dpk;         // (pk at the top interface is imposed on read, :1689-1693)
      acc[3] += r * dgz;
      if (p.nonhydro) {
        const double qa = top ? 0.0 : x.in(4, 0, 0, 0), qb = top ? 0.0 : x.in(4, ex, ey, 0), qc = x.in(4, 0, 0, 1), qd = x.in(4, ex, ey, 1);
        const double dsum = x.in(5, 0, 0) + x.in(5, ex, ey);
        const double s = p.dt * w / dsum;
        acc[3] += s * (role == 0 ? (qc - qb) : role == 1 ? -(qd - qa) : role == 2 ? (qd - qa) : -(qc - qb));
        if (!(top && role < 2)) acc[4] += s * (role == 0 ? -A : role == 1 ? -C : role == 2 ? C : A);
        if (role < 2) acc[5] += -s * (A * (qd - qa) + C * (qc - qb)) / dsum;     // the two dp of the edge share the layer of the output
      }
    };
#pragma unroll
    for (int role = 0; role < 4; role++) { side(true, role); side(false, role); }
  }
};

// split_p_grad (model/dyn_core_nlm.F90:1531-1643, TL model_tlmadm/dyn_core_tlm.F90:3592-3757) and grad1_p_update (:1781-1872, TL :4163-4293;
// d_ext = 0, so divg2 = 0 :726): the pressure gradient with the hydrostatic part of the previous acoustic sub-step blended in,
//   u <- u + beta du_prev ;  du = dt (A B + C D) / den ;  u <- (u + alpha du [+ dn]) rdx,   alpha = 1 - beta
// beta is the caller's beta_d (0 on the first sub-step, :373-375, where du_prev is not read).  The geometry is S_gradp's.
// ext (hydrostatic, d_ext > 0): the external-mode damping term divg2(i,j) - divg2(i+1,j) joins the wind before the metric factor
// (one_grad_p :1713-1727, :1758-1771 when grad1 = 0; grad1_p_update :1858, :1867 when grad1 = 1); divg2 is input 8 (the same
// value on every level, S_divg2).
// in: u v pkb gzb ppb dpb du_prev dv_prev divg2 ; out: u_new v_new du dv
struct S_gradp_beta {
  static constexpr int NI = 9, NO = 4;
  struct P { double dt, top, beta; int nonhydro, first, ext, grad1; };
  static constexpr int NT = 28;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0},
                                   {2, 0, 0, 0}, {2, 1, 0, 0}, {2, 0, 1, 0}, {2, 0, 0, 1}, {2, 1, 0, 1}, {2, 0, 1, 1},
                                   {3, 0, 0, 0}, {3, 1, 0, 0}, {3, 0, 1, 0}, {3, 0, 0, 1}, {3, 1, 0, 1}, {3, 0, 1, 1},
                                   {4, 0, 0, 0}, {4, 1, 0, 0}, {4, 0, 1, 0}, {4, 0, 0, 1}, {4, 1, 0, 1}, {4, 0, 1, 1},
                                   {5, 0, 0, 0}, {5, 1, 0, 0}, {5, 0, 1, 0}, {6, 0, 0, 0}, {7, 0, 0, 0},
                                   {8, 0, 0, 0}, {8, 1, 0, 0}, {8, 0, 1, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    const double alpha = 1.0 - p.beta;
    auto PK = [&](int di, int dj, int dk) -> T { return (x.kk + dk == 0) ? T(p.top) : x.in(2, di, dj, dk); };
    auto PP = [&](int di, int dj, int dk) -> T { return (x.kk + dk == 0) ? T(0.0) : x.in(4, di, dj, dk); };
    auto GZ = [&](int di, int dj, int dk) -> T { return x.in(3, di, dj, dk); };
    auto wk = [&](int di, int dj) -> T { return PK(di, dj, 1) - PK(di, dj, 0); };
    if (x.in_rect(g.is, g.ie, g.js, g.je + 1)) {
      T ub = x.in(0);
      if (!p.first) ub = ub + p.beta * x.in(6);
      if (p.ext) { if (p.grad1) ub = (ub + x.in(8)) - x.in(8, 1, 0); else ub = (x.in(8) - x.in(8, 1, 0)) + ub; }
      T du = p.dt / (wk(0, 0) + wk(1, 0)) * ((GZ(0, 0, 1) - GZ(1, 0, 0)) * (PK(1, 0, 1) - PK(0, 0, 0)) + (GZ(0, 0, 0) - GZ(1, 0, 1)) * (PK(0, 0, 1) - PK(1, 0, 0)));
      x.out(2, du);
      if (p.nonhydro) {
        T dn = p.dt / (x.in(5, 0, 0) + x.in(5, 1, 0)) * ((GZ(0, 0, 1) - GZ(1, 0, 0)) * (PP(1, 0, 1) - PP(0, 0, 0)) + (GZ(0, 0, 0) - GZ(1, 0, 1)) * (PP(0, 0, 1) - PP(1, 0, 0)));
        x.out(0, (ub + alpha * du + dn) * x.M(x.m.rdx));
      } else {
        x.out(0, (ub + alpha * du) * x.M(x.m.rdx));
      }
    }
    if (x.in_rect(g.is, g.ie + 1, g.js, g.je)) {
      T vb = x.in(1);
      if (!p.first) vb = vb + p.beta * x.in(7);
      if (p.ext) { if (p.grad1) vb = (vb + x.in(8)) - x.in(8, 0, 1); else vb = (x.in(8) - x.in(8, 0, 1)) + vb; }
      T dv = p.dt / (wk(0, 0) + wk(0, 1)) * ((GZ(0, 0, 1) - GZ(0, 1, 0)) * (PK(0, 1, 1) - PK(0, 0, 0)) + (GZ(0, 0, 0) - GZ(0, 1, 1)) * (PK(0, 0, 1) - PK(0, 1, 0)));
      x.out(3, dv);
      if (p.nonhydro) {
        T dn = p.dt / (x.in(5, 0, 0) + x.in(5, 0, 1)) * ((GZ(0, 0, 1) - GZ(0, 1, 0)) * (PP(0, 1, 1) - PP(0, 0, 0)) + (GZ(0, 0, 0) - GZ(0, 1, 1)) * (PP(0, 0, 1) - PP(0, 1, 0)));
        x.out(1, (vb + alpha * dv + dn) * x.M(x.m.rdy));
      } else {
        x.out(1, (vb + alpha * dv) * x.M(x.m.rdy));
      }
    }
  }
};

// a2b_ord2 (model/a2b_edge_nlm.F90:677-798, TL model_tlmadm/a2b_edge_tlm.F90 A2B_ORD2_TLM): A-grid -> corners by the four-cell mean, on
// the cube edges the mean of the two cells across the edge interpolated with edge_w/e/s/n, at the cube vertices the mean of the three
// cells that exist.  in: qin ; out: qout on is..ie+1, js..je+1
struct S_a2b_ord2 {
  static constexpr int NI = 1, NO = 1;
  struct P { int dummy; };
  static constexpr int NT = 4;
  static constexpr Tap taps[NT] = {{0, -1, -1, 0}, {0, 0, -1, 0}, {0, -1, 0, 0}, {0, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P&) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie + 1, g.js, g.je + 1)) return;
    const int i = x.i, j = x.j, npx = g.npx, npy = g.npy;
    const bool ei = (i == 1 || i == npx), ej = (j == 1 || j == npy);
    const double r3 = 1.0 / 3.0;
    T r;
    if (ei && ej) {
      if (i == 1 && j == 1) r = r3 * (x.in(0, 0, 0) + x.in(0, 0, -1) + x.in(0, -1, 0));
      else if (i == npx && j == 1) r = r3 * (x.in(0, -1, 0) + x.in(0, -1, -1) + x.in(0, 0, 0));
      else if (i == npx && j == npy) r = r3 * (x.in(0, -1, -1) + x.in(0, 0, -1) + x.in(0, -1, 0));
      else r = r3 * (x.in(0, 0, -1) + x.in(0, -1, -1) + x.in(0, 0, 0));
    } else if (ei) {
      const double e = x.M1(i == 1 ? x.m.edge_w : x.m.edge_e, j);
      r = e * (0.5 * (x.in(0, -1, -1) + x.in(0, 0, -1))) + (1.0 - e) * (0.5 * (x.in(0, -1, 0) + x.in(0, 0, 0)));
    } else if (ej) {
      const double e = x.M1(j == 1 ? x.m.edge_s : x.m.edge_n, i);
      r = e * (0.5 * (x.in(0, -1, -1) + x.in(0, -1, 0))) + (1.0 - e) * (0.5 * (x.in(0, 0, -1) + x.in(0, 0, 0)));
    } else {
      r = 0.25 * (x.in(0, -1, -1) + x.in(0, 0, -1) + x.in(0, -1, 0) + x.in(0, 0, 0));
    }
    x.out(0, r);
  }
};

// external-mode divergence (model/dyn_core_nlm.F90:707-724): the delp-weighted column mean of the corner divergence,
//   divg2 = c sum_k(dpc_k vt_k) / sum_k(dpc_k),  c = d_ext da_min_c;  written to every level of the output.
// in: dpc vt ; out: divg2 (K levels, all equal)
struct S_divg2 {
  static constexpr int NI = 2, NO = 1;
  struct P { int K; double c; };
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie + 1, g.js, g.je + 1)) return;
    T wk = x.in(0, 0), d = wk * x.in(1, 0);
    for (int k = 1; k < p.K; k++) { wk = wk + x.in(0, k); d = d + x.in(0, k) * x.in(1, k); }
    d = p.c * d / wk;
    for (int k = 0; k < p.K; k++) x.out(0, k, d);
  }
  template <class X> DEV static void eval_ad(X& x, const P& p) {
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie + 1, g.js, g.je + 1)) return;
    double a = 0.0, wk = 0.0, s = 0.0;
    for (int k = 0; k < p.K; k++) { a += x.oad(0, k); wk += x.in(0, k); s += x.in(0, k) * x.in(1, k); }
    const double f = a * p.c / wk, mean = s / wk;
    for (int k = 0; k < p.K; k++) { x.add(0, k, f * (x.in(1, k) - mean)); x.add(1, k, f * x.in(0, k)); }
  }
};

// out = a + b on a rectangle (flux / Courant-number accumulators, dyn_core/d_sw :913-931)
struct S_add2 {
  static constexpr int NI = 2, NO = 1;
  struct P { int i0, i1, j0, j1; };
  static constexpr int NT = 2;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    if (!x.in_rect(p.i0, p.i1, p.j0, p.j1)) return;
    x.out(0, x.in(0) + x.in(1));
  }
};

// ---------------------------------------------------------------------------------
// del2_cubed (model/dyn_core_nlm.F90:2090-2199; TL model_tlmadm/dyn_core_tlm.F90 DEL2_CUBED_TLM :4909): up to three
// passes of a 5-point Laplacian filter; pass n works on the compute domain widened by nt = ntimes - n cells.
// ---------------------------------------------------------------------------------
// corner averaging (:2146-2165): the three cells around each cube vertex take their mean (q(1,1)+q(0,1)+q(1,0))/3, ... ;
// every other cell of the region the pass reads is copied.   in: q ; out: q'
struct S_d2c_corner {
  static constexpr int NI = 1, NO = 1;
  struct P { int nt; };
  static constexpr int NT = 9;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {0, -1, 0, 0}, {0, 1, 0, 0}, {0, 0, -1, 0}, {0, 0, 1, 0}, {0, 1, -1, 0}, {0, -1, 1, 0}, {0, -1, -1, 0}, {0, 1, 1, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    const int h = p.nt + 1;
    if (!x.in_rect(g.is - h, g.ie + h, g.js - h, g.je + h)) return;
    const int i = x.i, j = x.j, ie = g.npx - 1, je = g.npy - 1, npx = g.npx, npy = g.npy;
    const double r3 = 1.0 / 3.0;
    // offsets of (A, B, C) = (interior corner cell, its x ghost, its y ghost) relative to this cell; the sum is (A + B) + C
    int ax = 0, ay = 0, bx = 0, by = 0, cx = 0, cy = 0; bool cor = true;
    if (i == 1 && j == 1)          { bx = -1; cy = -1; }
    else if (i == 0 && j == 1)     { ax = 1; cx = 1; cy = -1; }
    else if (i == 1 && j == 0)     { ay = 1; bx = -1; by = 1; }
    else if (i == ie && j == 1)    { bx = 1; cy = -1; }
    else if (i == npx && j == 1)   { ax = -1; cx = -1; cy = -1; }
    else if (i == ie && j == 0)    { ay = 1; bx = 1; by = 1; }
    else if (i == ie && j == je)   { bx = 1; cy = 1; }
    else if (i == npx && j == je)  { ax = -1; cx = -1; cy = 1; }
    else if (i == ie && j == npy)  { ay = -1; bx = 1; by = -1; }
    else if (i == 1 && j == je)    { bx = -1; cy = 1; }
    else if (i == 0 && j == je)    { ax = 1; cx = 1; cy = 1; }
    else if (i == 1 && j == npy)   { ay = -1; bx = -1; by = -1; }
    else cor = false;
    if (cor) { T a = x.in(0, ax, ay), b = x.in(0, bx, by), c = x.in(0, cx, cy); x.out(0, (a + b + c) * r3); }
    else x.out(0, x.in(0));
  }
};
// q <- q + cd rarea (fx(i) - fx(i+1) + fy(j) - fy(j+1))  on the domain widened by nt (:2189-2193).  in: q fx fy ; out: q'
struct S_d2c_upd {
  static constexpr int NI = 3, NO = 1;
  struct P { int nt; double cd; };
  static constexpr int NT = 5;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {1, 1, 0, 0}, {2, 0, 0, 0}, {2, 0, 1, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    const Geom& g = x.g;
    if (!x.in_rect(g.is - p.nt, g.ie + p.nt, g.js - p.nt, g.je + p.nt)) return;
    x.out(0, x.in(0) + p.cd * x.M(x.m.rarea) * (x.in(1) - x.in(1, 1, 0) + x.in(2) - x.in(2, 0, 1)));
  }
};

// dissipative heating added to pt on the top n_con layers (model/dyn_core_nlm.F90:1052-1099; pt is cp * virtual temperature / pkz).
//   hydrostatic: layers 1, 2: pt += hs / (cp delp pkz) ; below: dtmp = hs / (cp delp), pt += sign(min(delt, |dtmp|), dtmp) / pkz
//   otherwise  : pkz = exp(k1k log(rdg delp / delz pt)), dtmp = hs / (cv delp), pt += sign(min(delt, |dtmp|), dtmp) / pkz
// in: pt hs delp pkz|delz ; out: pt'
struct S_heat_pt {
  static constexpr int NI = 4, NO = 1;
  struct P { int n_con, hydrostatic; double cp_air, cv_air, delt, rdg, k1k; };
  static constexpr int NT = 4;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, 0, 0, 0}, {3, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie, g.js, g.je)) return;
    T pt = x.in(0);
    if (x.kk >= p.n_con) { x.out(0, pt); return; }
    T hs = x.in(1), dp = x.in(2);
    if (p.hydrostatic && x.kk < 2) { x.out(0, pt + hs / (p.cp_air * dp * x.in(3))); return; }
    T pkz = p.hydrostatic ? x.in(3) : m_exp(p.k1k * m_log(p.rdg * dp / x.in(3) * pt));
    T dtmp = hs / ((p.hydrostatic ? p.cp_air : p.cv_air) * dp);
    T mn = m_min(T(p.delt), m_abs(dtmp));
    x.out(0, pt + (val(dtmp) >= 0.0 ? mn : T(0.0) - mn) / pkz);
  }
};

struct DynParams;  // dyn.cu

}  // namespace fv3lm
