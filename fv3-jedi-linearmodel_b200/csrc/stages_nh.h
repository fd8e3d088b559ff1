// Non-hydrostatic stages: update_dz_c / update_dz_d (model/nh_utils_nlm.F90:43-296), edge_profile
// (:1519-1625), the semi-implicit vertical solver SIM1_solver (:1177-1308) behind Riem_Solver_c
// (:297-401) and Riem_Solver3 (model/nh_core_nlm.F90:40-206).
// TL: model_tlmadm/nh_utils_tlm.F90 (UPDATE_DZ_C_TLM :51, UPDATE_DZ_D_TLM :381, RIEM_SOLVER_C_TLM :723,
// SIM1_SOLVER_TLM :2548, EDGE_PROFILE_TLM :3319), nh_core_tlm.F90 :49; AD: nh_utils_adm.F90, nh_core_adm.F90.
// Column solvers: one thread per column; NL/TL share a templated sweep, the adjoint is hand-written.
#pragma once
#include "engine.h"
#include "mosaic.h"
#include "stages_dsw.h"
#include "stages_remap.h"

namespace fv3lm {

constexpr double DZ_MIN = 2.0;   // model/nh_utils_nlm.F90:38

// gz(K) = zs ; gz(k) = gz(k+1) - delz(k)   (dyn_core_nlm.F90:341-351).  in: delz zs ; out: gz
struct S_gz_init {
  static constexpr int NI = 2, NO = 1;
  struct P { int K; };
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie, g.js, g.je)) return;
    T z = x.in(1, 0);
    x.out(0, p.K, z);
    for (int k = p.K - 1; k >= 0; k--) { z = z - x.in(0, k); x.out(0, k, z); }
  }
  template <class X> DEV static void eval_ad(X& x, const P& p) {
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie, g.js, g.je)) return;
    double a = 0.0;
    for (int k = 0; k < p.K; k++) { a += x.oad(0, k); x.add(0, k, -a); }
    a += x.oad(0, p.K);
    x.add(1, 0, a);
  }
};

// interface winds of update_dz_c (:68-98).  in: ut vt ; out: xfx yfx  (K+1 levels)
struct S_dzc_wind {
  static constexpr int NI = 2, NO = 2;
  struct P { LevD dp0; int K; };
  static constexpr int NT = 8;
  static constexpr Tap taps[NT] = {{0, 0, 0, -2}, {0, 0, 0, -1}, {0, 0, 0, 0}, {0, 0, 0, 1}, {1, 0, 0, -2}, {1, 0, 0, -1}, {1, 0, 0, 0}, {1, 0, 0, 1}};
  template <class X> DEV static typename X::T iface(const X& x, const P& p, int f) {
    const int k = x.kk, K = p.K;
    if (k == 0) return x.in(f, 0, 0, 0) + (x.in(f, 0, 0, 0) - x.in(f, 0, 0, 1)) * (p.dp0.v[0] / (p.dp0.v[0] + p.dp0.v[1]));
    if (k == K) return x.in(f, 0, 0, -1) + (x.in(f, 0, 0, -1) - x.in(f, 0, 0, -2)) * (p.dp0.v[K - 1] / (p.dp0.v[K - 2] + p.dp0.v[K - 1]));
    return (p.dp0.v[k] * x.in(f, 0, 0, -1) + p.dp0.v[k - 1] * x.in(f, 0, 0, 0)) * (1.0 / (p.dp0.v[k - 1] + p.dp0.v[k]));
  }
  template <class X> DEV static void eval(X& x, const P& p) {
    const Geom& g = x.g;
    if (x.in_rect(g.is - 1, g.ie + 2, g.js - 1, g.je + 1)) x.out(0, iface(x, p, 0));
    if (x.in_rect(g.is - 1, g.ie + 1, g.js - 1, g.je + 2)) x.out(1, iface(x, p, 1));
  }
};
// upwind flux of gz (:105-126).  DIR 0: in: xfx gz ; out: fx
template <int DIR> struct S_dzc_flux {
  static constexpr int NI = 2, NO = 1;
  struct P { int dummy; };
  static constexpr int NT = 3;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {1, DIR == 0 ? -1 : 0, DIR == 0 ? 0 : -1, 0}};
  template <class X> DEV static void eval(X& x, const P&) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (DIR == 0) { if (!x.in_rect(g.is - 1, g.ie + 2, g.js - 1, g.je + 1)) return; }
    else { if (!x.in_rect(g.is - 1, g.ie + 1, g.js - 1, g.je + 2)) return; }
    T c = x.in(0);
    x.out(0, c * (val(c) > 0.0 ? (DIR == 0 ? x.in(1, -1, 0) : x.in(1, 0, -1)) : x.in(1)));
  }
};
// gz update (:127-133).  in: gz fx fy xfx yfx ; out: gzn
struct S_dzc_upd {
  static constexpr int NI = 5, NO = 1;
  struct P { int dummy; };
  static constexpr int NT = 9;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {1, 1, 0, 0}, {2, 0, 0, 0}, {2, 0, 1, 0}, {3, 0, 0, 0}, {3, 1, 0, 0}, {4, 0, 0, 0}, {4, 0, 1, 0}};
  template <class X> DEV static void eval(X& x, const P&) {
    const Geom& g = x.g;
    if (!x.in_rect(g.is - 1, g.ie + 1, g.js - 1, g.je + 1)) return;
    double ar = x.M(x.m.area);
    x.out(0, (x.in(0) * ar + (x.in(1) - x.in(1, 1, 0)) + (x.in(2) - x.in(2, 0, 1))) /
                 (ar + (x.in(3) - x.in(3, 1, 0)) + (x.in(4) - x.in(4, 0, 1))));
  }
};
// ws and the monotonicity clamp (:136-145, :284-293).  in: zn zs ; out: z ws
struct S_dz_clamp {
  static constexpr int NI = 2, NO = 2;
  struct P { int K; double rdt; int halo; };
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is - p.halo, g.ie + p.halo, g.js - p.halo, g.je + p.halo)) return;
    T zb = x.in(0, p.K);
    x.out(1, 0, (x.in(1, 0) - zb) * p.rdt);
    x.out(0, p.K, zb);
    for (int k = p.K - 1; k >= 0; k--) { zb = m_max(x.in(0, k), zb + DZ_MIN); x.out(0, k, zb); }
  }
  template <class X> DEV static void eval_ad(X& x, const P& p) {
    const Geom& g = x.g;
    if (!x.in_rect(g.is - p.halo, g.ie + p.halo, g.js - p.halo, g.je + p.halo)) return;
    double carry = 0.0;
    for (int k = 0; k < p.K; k++) {
      double a = x.oad(0, k) + carry;
      if (x.in(0, k) > x.outv(0, k + 1) + DZ_MIN) { x.add(0, k, a); carry = 0.0; }
      else carry = a;
    }
    double a = x.oad(0, p.K) + carry;
    double w = x.oad(1, 0) * p.rdt;
    x.add(0, p.K, a - w);
    x.add(1, 0, w);
  }
};

// edge_profile (non-uniform branch, limiter = 0): linear in q, coefficients from dp0 only.  in: q ; out: qe (K+1)
struct S_edge_profile {
  static constexpr int NI = 1, NO = 1;
  // The elimination coefficients depend on the reference thicknesses dp0 only -- the same for every column -- so they are computed once
  // on the host (make) and the kernels keep no per-column work array: the forward sweep parks its result in the output array and the
  // back substitution updates it in place (ncu, round 2: the version with per-thread qe[] / gam[] arrays ran at 0.9 TB/s of
  // algorithmic traffic, its local arrays spilling through L2 into HBM).
  struct P { int K; int i0, i1, j0, j1; double xt1_0, xt1_b, a_bot, xt2; LevD gk, bet, gam; };
  static P make(const LevD& dp0, int K, int i0, int i1, int j0, int j1) {
    P p; p.K = K; p.i0 = i0; p.i1 = i1; p.j0 = j0; p.j1 = j1;
    for (int k = 0; k < 96; k++) p.gk.v[k] = p.bet.v[k] = p.gam.v[k] = 0.0;
    const double g0 = dp0.v[1] / dp0.v[0];
    p.xt1_0 = 2.0 * g0 * (g0 + 1.0); p.bet.v[0] = g0 * (g0 + 0.5);
    p.gam.v[0] = (1.0 + g0 * (g0 + 1.5)) / p.bet.v[0];
    double gk = 0.0;
    for (int k = 1; k < K; k++) {
      gk = dp0.v[k - 1] / dp0.v[k];
      p.gk.v[k] = gk;
      p.bet.v[k] = 2.0 + 2.0 * gk - p.gam.v[k - 1];
      p.gam.v[k] = gk / p.bet.v[k];
    }
    p.a_bot = 1.0 + gk * (gk + 1.5);
    p.xt1_b = 2.0 * gk * (gk + 1.0);
    p.xt2 = gk * (gk + 0.5) - p.a_bot * p.gam.v[K - 1];
    return p;
  }
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    if (!x.in_rect(p.i0, p.i1, p.j0, p.j1)) return;
    const int K = p.K;
    T qm = x.in(0, 0), q0 = x.in(0, 1);                 // Q(k-1), Q(k)
    T qe = (p.xt1_0 * qm + q0) / p.bet.v[0];
    x.out(0, 0, qe);
    T qmm = qm;                                         // Q(k-2)
#pragma unroll 4
    for (int k = 1; k < K; k++) {
      if (k > 1) { qmm = qm; qm = q0; q0 = x.in(0, k); }
      qe = (3.0 * (qm + p.gk.v[k] * q0) - qe) / p.bet.v[k];
      x.out(0, k, qe);
    }
    // here q0 = Q(K-1), qm = Q(K-2)
    T nx = (p.xt1_b * q0 + qm - p.a_bot * qe) / p.xt2;
    x.out(0, K, nx);
    (void)qmm;
    for (int k = K - 1; k >= 0; k--) { nx = x.rd(0, k) - p.gam.v[k] * nx; x.out(0, k, nx); }
  }
  // the operator is linear with constant coefficients: transpose the elimination.  The adjoint of the output (dead after this op) is
  // the work array of the transposed back substitution.
  template <class X> DEV static void eval_ad(X& x, const P& p) {
    if (!x.in_rect(p.i0, p.i1, p.j0, p.j1)) return;
    const int K = p.K;
    double a = x.oad(0, 0);
    for (int k = 0; k < K; k++) {                       // ad[k+1] -= gam[k] * ad[k]
      const double upd = -(p.gam.v[k] * a);
      a = x.oad(0, k + 1) + upd;
      x.oad_add(0, k + 1, upd);
    }
    // a = ad[K]
    double n = a / p.xt2;
    double q_hi = p.xt1_b * n;                          // pending contribution to q_ad[k] of the row above (k = K-1 now)
    double q_lo = n;                                    // ... and to q_ad[k-1]
    double carry = -(p.a_bot * n);                      // ad[K-1] -= a_bot n
    for (int k = K - 1; k >= 1; k -= 4) {        // four levels per batch: work array and accumulators requested together
      double o[4], t[4];
#pragma unroll
      for (int u = 0; u < 4; u++) if (k - u >= 1) { o[u] = x.oad(0, k - u); t[u] = x.iad(0, k - u); }
#pragma unroll
      for (int u = 0; u < 4; u++) if (k - u >= 1) {
        n = (o[u] + carry) / p.bet.v[k - u];
        x.iad_set(0, k - u, t[u] + (q_hi + 3.0 * p.gk.v[k - u] * n));
        q_hi = q_lo + 3.0 * n; q_lo = 0.0;
        carry = -n;
      }
    }
    n = (x.oad(0, 0) + carry) / p.bet.v[0];
    // K >= 2: q_hi holds what rows 1.. gave to q_ad[0]; row 0 adds xt1_0 n to q_ad[0] and n to q_ad[1]
    x.add(0, 0, q_hi + p.xt1_0 * n);
    x.add(0, 1, n);
  }
};

// zh update of update_dz_d (:258-281).  in: zh fx fy ra_x ra_y fx2 fy2 ; out: zhn   (K+1 levels)
struct S_dzd_upd {
  static constexpr int NI = 7, NO = 1;
  struct P { LevD on; };
  static constexpr int NT = 11;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {1, 1, 0, 0}, {2, 0, 0, 0}, {2, 0, 1, 0}, {3, 0, 0, 0}, {4, 0, 0, 0},
                                   {5, 0, 0, 0}, {5, 1, 0, 0}, {6, 0, 0, 0}, {6, 0, 1, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is, g.ie, g.js, g.je)) return;
    double ar = x.M(x.m.area);
    T r = (x.in(0) * ar + (x.in(1) - x.in(1, 1, 0)) + (x.in(2) - x.in(2, 0, 1))) / ((x.in(3) + x.in(4)) - ar);
    if (p.on.v[x.kk] != 0.0) r = r + ((x.in(5) - x.in(5, 1, 0)) + (x.in(6) - x.in(6, 0, 1))) * x.M(x.m.rarea);
    x.out(0, r);
  }
};
// out = a * c   (gz = zh * grav, dyn_core_nlm.F90:823-829)
struct S_scale {
  static constexpr int NI = 1, NO = 1;
  struct P { double c; int halo; };
  static constexpr int NT = 1;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    const Geom& g = x.g;
    if (!x.in_rect(g.is - p.halo, g.ie + p.halo, g.js - p.halo, g.je + p.halo)) return;
    x.out(0, p.c * x.in(0));
  }
};

// ---------------------------------------------------------------------------------
// Riem_Solver_c (mode 0) / Riem_Solver3 (mode 1) around SIM1_solver
//   mode 0: in: delp pt gz w3 ws hs ; out: pef gz_new  (w unused)          halo 1
//   mode 1: in: delp pt zh w  ws zs ; out: ppe zh_new w_new delz           halo 0
// ---------------------------------------------------------------------------------
struct S_riem {
  static constexpr int NI = 6, NO = 4;
  struct P { int K, mode, halo; double dt, akap, ptop, rdgas, grav, p_fac; };
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is - p.halo, g.ie + p.halo, g.js - p.halo, g.je + p.halo)) return;
    const int K = p.K;
    const double gama = 1.0 / (1.0 - p.akap), rgrav = 1.0 / p.grav, t1g = gama * 2.0 * p.dt * p.dt, rdt = 1.0 / p.dt, capa1 = p.akap - 1.0;
    // work arrays are local memory streamed through HBM: cheap quantities (dm, dz, g, bb, dd) are recomputed from the inputs
    T pm2[KMAX], pem[KMAX + 1], pe[KMAX], pp[KMAX + 1], gam[KMAX], aa[KMAX], w2[KMAX], pe2[KMAX + 1], dz[KMAX];
    auto DM = [&](int k) -> T { return x.in(0, k) * rgrav; };
    auto DZ = [&](int k) -> T { return x.in(2, k + 1) - x.in(2, k); };
    auto G_ = [&](int k) -> T { return x.in(0, k) / x.in(0, k + 1); };
    pem[0] = T(p.ptop);
    for (int k = 0; k < K; k++) pem[k + 1] = pem[k] + x.in(0, k);
#pragma unroll 4
    for (int k = 0; k < K; k++) {
      T dp = x.in(0, k);
      pm2[k] = (p.mode == 0) ? dp / m_log(pem[k + 1] / pem[k]) : dp / (m_log(pem[k + 1]) - m_log(pem[k]));
    }
    T ws = x.in(4, 0);
    // ---- SIM1_solver
#pragma unroll 4
    for (int k = 0; k < K; k++) pe[k] = m_exp(gama * m_log(-DM(k) / DZ(k) * p.rdgas * x.in(1, k))) - pm2[k];
    auto BB = [&](int k) -> T { return k < K - 1 ? 2.0 * (1.0 + G_(k)) : T(2.0); };
    auto DD = [&](int k) -> T { return k < K - 1 ? 3.0 * (pe[k] + G_(k) * pe[k + 1]) : 3.0 * pe[K - 1]; };
    T bet = BB(0);
    pp[0] = T(0.0); pp[1] = DD(0) / bet;
    for (int k = 1; k < K; k++) { gam[k] = G_(k - 1) / bet; bet = BB(k) - gam[k]; pp[k + 1] = (DD(k) - pp[k]) / bet; }
    for (int k = K - 1; k >= 1; k--) pp[k] = pp[k] - gam[k] * pp[k + 1];
#pragma unroll 4
    for (int k = 1; k < K; k++) aa[k] = t1g / (DZ(k - 1) + DZ(k)) * (pem[k] + pp[k]);
    bet = DM(0) - aa[1];
    w2[0] = (DM(0) * x.in(3, 0) + p.dt * pp[1]) / bet;
    for (int k = 1; k < K - 1; k++) {
      gam[k] = aa[k] / bet;
      bet = DM(k) - (aa[k] + aa[k + 1] + aa[k] * gam[k]);
      w2[k] = (DM(k) * x.in(3, k) + p.dt * (pp[k + 1] - pp[k]) - aa[k] * w2[k - 1]) / bet;
    }
    T p1 = t1g / DZ(K - 1) * (pem[K] + pp[K]);
    gam[K - 1] = aa[K - 1] / bet;
    bet = DM(K - 1) - (aa[K - 1] + p1 + aa[K - 1] * gam[K - 1]);
    w2[K - 1] = (DM(K - 1) * x.in(3, K - 1) + p.dt * (pp[K] - pp[K - 1]) - p1 * ws - aa[K - 1] * w2[K - 2]) / bet;
    for (int k = K - 2; k >= 0; k--) w2[k] = w2[k] - gam[k + 1] * w2[k + 1];
    pe2[0] = T(0.0);
    for (int k = 0; k < K; k++) pe2[k + 1] = pe2[k] + DM(k) * (w2[k] - x.in(3, k)) * rdt;
    p1 = (pe2[K - 1] + 2.0 * pe2[K]) * rmp::r3;
    dz[K - 1] = -DM(K - 1) * p.rdgas * x.in(1, K - 1) * m_exp(capa1 * m_log(m_max(p.p_fac * pm2[K - 1], p1 + pm2[K - 1])));
    for (int k = K - 2; k >= 0; k--) {
      p1 = (pe2[k] + BB(k) * pe2[k + 1] + G_(k) * pe2[k + 2]) * rmp::r3 - G_(k) * p1;
      dz[k] = -DM(k) * p.rdgas * x.in(1, k) * m_exp(capa1 * m_log(m_max(p.p_fac * pm2[k], p1 + pm2[k])));
    }
    // ---- outputs
    if (p.mode == 0) {
      x.out(0, 0, T(p.ptop));
      for (int k = 1; k <= K; k++) x.out(0, k, pe2[k] + pem[k]);
      T z = x.in(5, 0);
      x.out(1, K, z);
      for (int k = K - 1; k >= 0; k--) { z = z - dz[k] * p.grav; x.out(1, k, z); }
    } else {
      for (int k = 0; k <= K; k++) x.out(0, k, pe2[k]);
      T z = x.in(5, 0);
      x.out(1, K, z);
      for (int k = K - 1; k >= 0; k--) { z = z - dz[k]; x.out(1, k, z); }
      for (int k = 0; k < K; k++) { x.out(2, k, w2[k]); x.out(3, k, dz[k]); }
    }
  }

  // Hand-written adjoint (forward recomputation in double keeping the elimination intermediates, then the reverse
  // sweep).  A leaner variant that recomputes dm/dz/g/bb/dd/gam from the inputs and propagates the Thomas
  // back-substitution adjoints in place (26 instead of 45 work arrays) measured SLOWER on B200 (69 vs 61 ms per 9
  // launches at C180: 116 registers and longer dependent chains), so the store-everything form is kept; see git history.
  template <class X> DEV static void eval_ad(X& x, const P& p) {
    const Geom& g = x.g;
    if (!x.in_rect(g.is - p.halo, g.ie + p.halo, g.js - p.halo, g.je + p.halo)) return;
    const int K = p.K;
    const double gama = 1.0 / (1.0 - p.akap), rgrav = 1.0 / p.grav, t1g = gama * 2.0 * p.dt * p.dt, rdt = 1.0 / p.dt, capa1 = p.akap - 1.0, r3 = rmp::r3;
    // ---- forward recomputation (double) keeping the elimination intermediates
    double dp[KMAX], dm[KMAX], pm2[KMAX], pem[KMAX + 1], dz[KMAX], pt[KMAX], w1[KMAX], pe[KMAX], g_[KMAX], bb[KMAX], dd[KMAX];
    double ppt[KMAX + 1], pp[KMAX + 1], gamC[KMAX], betC[KMAX], aa[KMAX + 1], gam2[KMAX], betE[KMAX], w2t[KMAX], w2[KMAX], pe2[KMAX + 1], p1v[KMAX], dzn[KMAX];
    pem[0] = p.ptop;
    for (int k = 0; k < K; k++) { dp[k] = x.in(0, k); pem[k + 1] = pem[k] + dp[k]; pt[k] = x.in(1, k); w1[k] = x.in(3, k); }
    for (int k = 0; k < K; k++) {
      dz[k] = x.in(2, k + 1) - x.in(2, k);
      pm2[k] = (p.mode == 0) ? dp[k] / log(pem[k + 1] / pem[k]) : dp[k] / (log(pem[k + 1]) - log(pem[k]));
      dm[k] = dp[k] * rgrav;
    }
    const double ws = x.in(4, 0);
    for (int k = 0; k < K; k++) pe[k] = exp(gama * log(-dm[k] / dz[k] * p.rdgas * pt[k])) - pm2[k];
    for (int k = 0; k < K - 1; k++) { g_[k] = dm[k] / dm[k + 1]; bb[k] = 2.0 * (1.0 + g_[k]); dd[k] = 3.0 * (pe[k] + g_[k] * pe[k + 1]); }
    bb[K - 1] = 2.0; dd[K - 1] = 3.0 * pe[K - 1]; g_[K - 1] = 0.0;
    betC[0] = bb[0]; ppt[0] = 0.0; ppt[1] = dd[0] / betC[0];
    for (int k = 1; k < K; k++) { gamC[k] = g_[k - 1] / betC[k - 1]; betC[k] = bb[k] - gamC[k]; ppt[k + 1] = (dd[k] - ppt[k]) / betC[k]; }
    pp[K] = ppt[K]; pp[0] = 0.0;
    for (int k = K - 1; k >= 1; k--) pp[k] = ppt[k] - gamC[k] * pp[k + 1];
    for (int k = 1; k < K; k++) aa[k] = t1g / (dz[k - 1] + dz[k]) * (pem[k] + pp[k]);
    betE[0] = dm[0] - aa[1];
    w2t[0] = (dm[0] * w1[0] + p.dt * pp[1]) / betE[0];
    for (int k = 1; k < K - 1; k++) {
      gam2[k] = aa[k] / betE[k - 1];
      betE[k] = dm[k] - (aa[k] + aa[k + 1] + aa[k] * gam2[k]);
      w2t[k] = (dm[k] * w1[k] + p.dt * (pp[k + 1] - pp[k]) - aa[k] * w2t[k - 1]) / betE[k];
    }
    const double p1b = t1g / dz[K - 1] * (pem[K] + pp[K]);
    gam2[K - 1] = aa[K - 1] / betE[K - 2];
    betE[K - 1] = dm[K - 1] - (aa[K - 1] + p1b + aa[K - 1] * gam2[K - 1]);
    w2t[K - 1] = (dm[K - 1] * w1[K - 1] + p.dt * (pp[K] - pp[K - 1]) - p1b * ws - aa[K - 1] * w2t[K - 2]) / betE[K - 1];
    w2[K - 1] = w2t[K - 1];
    for (int k = K - 2; k >= 0; k--) w2[k] = w2t[k] - gam2[k + 1] * w2[k + 1];
    pe2[0] = 0.0;
    for (int k = 0; k < K; k++) pe2[k + 1] = pe2[k] + dm[k] * (w2[k] - w1[k]) * rdt;
    p1v[K - 1] = (pe2[K - 1] + 2.0 * pe2[K]) * r3;
    for (int k = K - 2; k >= 0; k--) p1v[k] = (pe2[k] + bb[k] * pe2[k + 1] + g_[k] * pe2[k + 2]) * r3 - g_[k] * p1v[k + 1];
    for (int k = 0; k < K; k++) dzn[k] = -dm[k] * p.rdgas * pt[k] * exp(capa1 * log(fmax(p.p_fac * pm2[k], p1v[k] + pm2[k])));
    // ---- adjoint seeds from the outputs
    double dm_ad[KMAX], pm2_ad[KMAX], pem_ad[KMAX + 1], dz_ad[KMAX], pt_ad[KMAX], w1_ad[KMAX], pe_ad[KMAX], g_ad[KMAX], bb_ad[KMAX], dd_ad[KMAX];
    double pp_ad[KMAX + 1], ppt_ad[KMAX + 1], gamC_ad[KMAX], betC_ad[KMAX], aa_ad[KMAX + 1], gam2_ad[KMAX], betE_ad[KMAX], w2_ad[KMAX], w2t_ad[KMAX], pe2_ad[KMAX + 1], dzn_ad[KMAX];
    for (int k = 0; k < K; k++) { dm_ad[k] = pm2_ad[k] = dz_ad[k] = pt_ad[k] = w1_ad[k] = pe_ad[k] = g_ad[k] = bb_ad[k] = dd_ad[k] = 0.0; gamC_ad[k] = betC_ad[k] = gam2_ad[k] = betE_ad[k] = w2_ad[k] = w2t_ad[k] = dzn_ad[k] = 0.0; aa_ad[k] = 0.0; }
    for (int k = 0; k <= K; k++) { pem_ad[k] = pp_ad[k] = ppt_ad[k] = pe2_ad[k] = 0.0; }
    aa_ad[K] = 0.0;
    double ws_ad = 0.0, zb_ad = 0.0;   // zb: bottom boundary height (hs or zs)
    {
      // height recurrence  z(K) = zb ; z(k) = z(k+1) - c dz(k)
      const double c = (p.mode == 0) ? p.grav : 1.0;
      double a = 0.0;
      for (int k = 0; k < K; k++) { a += x.oad(1, k); dzn_ad[k] -= a * c; }
      a += x.oad(1, K);
      zb_ad = a;
      if (p.mode == 0) { for (int k = 1; k <= K; k++) { double o = x.oad(0, k); pe2_ad[k] += o; pem_ad[k] += o; } }
      else { for (int k = 0; k <= K; k++) pe2_ad[k] += x.oad(0, k); for (int k = 0; k < K; k++) { w2_ad[k] += x.oad(2, k); dzn_ad[k] += x.oad(3, k); } }
    }
    // ---- G: dz_new and the p1 recursion
    {
      double p1_ad = 0.0;   // adjoint of p1v[k], carried downwards (k increasing)
      for (int k = 0; k < K; k++) {
        const double M = fmax(p.p_fac * pm2[k], p1v[k] + pm2[k]);
        const double a = dzn_ad[k];
        dm_ad[k] += a * dzn[k] / dm[k]; pt_ad[k] += a * dzn[k] / pt[k];
        const double M_ad = a * dzn[k] * capa1 / M;
        double P = p1_ad;
        if (p.p_fac * pm2[k] > p1v[k] + pm2[k]) pm2_ad[k] += M_ad * p.p_fac;
        else { P += M_ad; pm2_ad[k] += M_ad; }
        if (k < K - 1) {
          pe2_ad[k] += P * r3; bb_ad[k] += P * r3 * pe2[k + 1]; pe2_ad[k + 1] += P * r3 * bb[k];
          g_ad[k] += P * (r3 * pe2[k + 2] - p1v[k + 1]); pe2_ad[k + 2] += P * r3 * g_[k];
          p1_ad = -g_[k] * P;
        } else {
          pe2_ad[K - 1] += P * r3; pe2_ad[K] += 2.0 * P * r3;
        }
      }
    }
    // ---- F: pe2 prefix sum
    for (int k = K - 1; k >= 0; k--) {
      const double a = pe2_ad[k + 1];
      pe2_ad[k] += a; dm_ad[k] += a * (w2[k] - w1[k]) * rdt; w2_ad[k] += a * dm[k] * rdt; w1_ad[k] -= a * dm[k] * rdt;
    }
    // ---- E: tridiagonal solve for w
    for (int k = 0; k <= K - 2; k++) { w2t_ad[k] += w2_ad[k]; gam2_ad[k + 1] -= w2_ad[k] * w2[k + 1]; w2_ad[k + 1] -= gam2[k + 1] * w2_ad[k]; }
    w2t_ad[K - 1] += w2_ad[K - 1];
    double p1b_ad = 0.0;
    {
      const int k = K - 1;
      const double n_ad = w2t_ad[k] / betE[k]; double b_ad = -w2t_ad[k] * w2t[k] / betE[k];
      dm_ad[k] += n_ad * w1[k]; w1_ad[k] += n_ad * dm[k]; pp_ad[K] += n_ad * p.dt; pp_ad[K - 1] -= n_ad * p.dt;
      p1b_ad -= n_ad * ws; ws_ad -= n_ad * p1b; aa_ad[k] -= n_ad * w2t[k - 1]; w2t_ad[k - 1] -= n_ad * aa[k];
      dm_ad[k] += b_ad; aa_ad[k] -= b_ad * (1.0 + gam2[k]); p1b_ad -= b_ad; gam2_ad[k] -= b_ad * aa[k];
      aa_ad[k] += gam2_ad[k] / betE[k - 1]; betE_ad[k - 1] -= gam2_ad[k] * gam2[k] / betE[k - 1];
      dz_ad[K - 1] -= p1b_ad * p1b / dz[K - 1]; pem_ad[K] += p1b_ad * t1g / dz[K - 1]; pp_ad[K] += p1b_ad * t1g / dz[K - 1];
    }
    for (int k = K - 2; k >= 1; k--) {
      const double n_ad = w2t_ad[k] / betE[k]; const double b_ad = betE_ad[k] - w2t_ad[k] * w2t[k] / betE[k];
      dm_ad[k] += n_ad * w1[k] + b_ad; w1_ad[k] += n_ad * dm[k]; pp_ad[k + 1] += n_ad * p.dt; pp_ad[k] -= n_ad * p.dt;
      aa_ad[k] -= n_ad * w2t[k - 1] + b_ad * (1.0 + gam2[k]); w2t_ad[k - 1] -= n_ad * aa[k]; aa_ad[k + 1] -= b_ad; gam2_ad[k] -= b_ad * aa[k];
      aa_ad[k] += gam2_ad[k] / betE[k - 1]; betE_ad[k - 1] -= gam2_ad[k] * gam2[k] / betE[k - 1];
    }
    {
      const double n_ad = w2t_ad[0] / betE[0]; const double b_ad = betE_ad[0] - w2t_ad[0] * w2t[0] / betE[0];
      dm_ad[0] += n_ad * w1[0] + b_ad; w1_ad[0] += n_ad * dm[0]; pp_ad[1] += n_ad * p.dt; aa_ad[1] -= b_ad;
    }
    // ---- D: aa
    for (int k = 1; k < K; k++) {
      const double a = aa_ad[k], s = dz[k - 1] + dz[k];
      dz_ad[k - 1] -= a * aa[k] / s; dz_ad[k] -= a * aa[k] / s; pem_ad[k] += a * t1g / s; pp_ad[k] += a * t1g / s;
    }
    // ---- C: tridiagonal solve for pp
    for (int k = 1; k <= K - 1; k++) { ppt_ad[k] += pp_ad[k]; gamC_ad[k] -= pp_ad[k] * pp[k + 1]; pp_ad[k + 1] -= gamC[k] * pp_ad[k]; }
    ppt_ad[K] += pp_ad[K];
    for (int k = K - 1; k >= 1; k--) {
      const double n_ad = ppt_ad[k + 1] / betC[k]; const double b_ad = betC_ad[k] - ppt_ad[k + 1] * ppt[k + 1] / betC[k];
      dd_ad[k] += n_ad; ppt_ad[k] -= n_ad; bb_ad[k] += b_ad; gamC_ad[k] -= b_ad;
      g_ad[k - 1] += gamC_ad[k] / betC[k - 1]; betC_ad[k - 1] -= gamC_ad[k] * gamC[k] / betC[k - 1];
    }
    dd_ad[0] += ppt_ad[1] / betC[0]; bb_ad[0] += betC_ad[0] - ppt_ad[1] * ppt[1] / betC[0];
    // ---- B
    for (int k = 0; k < K - 1; k++) {
      pe_ad[k] += 3.0 * dd_ad[k]; g_ad[k] += 3.0 * dd_ad[k] * pe[k + 1] + 2.0 * bb_ad[k]; pe_ad[k + 1] += 3.0 * dd_ad[k] * g_[k];
      dm_ad[k] += g_ad[k] / dm[k + 1]; dm_ad[k + 1] -= g_ad[k] * g_[k] / dm[k + 1];
    }
    pe_ad[K - 1] += 3.0 * dd_ad[K - 1];
    // ---- A
    for (int k = 0; k < K; k++) {
      const double a = pe_ad[k], Ek = pe[k] + pm2[k];
      pm2_ad[k] -= a; dm_ad[k] += a * Ek * gama / dm[k]; dz_ad[k] -= a * Ek * gama / dz[k]; pt_ad[k] += a * Ek * gama / pt[k];
    }
    // ---- wrappers: dm = dp rgrav ; pm2 = dp / (log pem(k+1) - log pem(k)) ; dz = z(k+1) - z(k) ; pem prefix sum
    double dp_ad[KMAX];
    for (int k = 0; k < K; k++) {
      const double dl = log(pem[k + 1]) - log(pem[k]);
      dp_ad[k] = dm_ad[k] * rgrav + pm2_ad[k] / dl;
      const double dl_ad = -pm2_ad[k] * pm2[k] / dl;
      pem_ad[k + 1] += dl_ad / pem[k + 1]; pem_ad[k] -= dl_ad / pem[k];
    }
    { double a = 0.0; for (int k = K; k >= 1; k--) { a += pem_ad[k]; dp_ad[k - 1] += a; } }
    for (int k = 0; k < K; k++) { x.add(0, k, dp_ad[k]); x.add(1, k, pt_ad[k]); x.add(3, k, w1_ad[k]); x.add(2, k + 1, dz_ad[k]); x.add(2, k, -dz_ad[k]); }
    x.add(4, 0, ws_ad);
    x.add(5, 0, zb_ad);
  }
};

}  // namespace fv3lm
