// Riem_Solver_c / Riem_Solver3 (SIM1_solver and SIM_solver, model/nh_utils_nlm.F90:297-401, :1177-1308, :1310-1466, model/nh_core_nlm.F90:40-206)
// decomposed into level-parallel stencil stages and thin column recurrences.
//
// The monolithic one-thread-per-column kernel (S_riem, stages_nh.h) keeps 13-45 work arrays of 72 levels in local
// memory: ncu (C180) shows 27 GB of DRAM traffic per adjoint launch for 1.8 GB of algorithmic bytes, 34 % occupancy
// and one column per thread -- it neither saturates one GPU nor scales when the cube is sharded.  Here everything
// that is independent across levels (the exp/log equation of state, the tridiagonal coefficients, the new layer
// thickness) runs as ordinary stencil stages with one thread per cell (72x more parallelism, coalesced, NL/TL/AD from
// the generic engine); only the true recurrences stay column kernels, each with one or two small work arrays:
//   S_cum      pem = ptop + cumsum(delp) ; z = zb - cumsum_from_below(c dz)
//   S_tri      Thomas solve of a general tridiagonal system; adjoint = solve with the transposed matrix
//   S_rs_pe2   pe2 = cumsum(dm (w2 - w1) / dt)
//   S_rs_p1    backward recurrence of the pressure-perturbation average
#pragma once
#include "engine.h"
#include "stages_remap.h"   // KMAX, rmp::r3

namespace fv3lm {

// ---- column: prefix / suffix sums ---------------------------------------------------------------------------
// mode 0: out(0) = top ; out(k+1) = out(k) + in0(k)                  in: a            ; out: s (K+1)
// mode 1: out(K) = in1 ; out(k) = out(k+1) - c in0(k)                in: a, bottom(2D) ; out: s (K+1)
struct S_cum {
  static constexpr int NI = 2, NO = 1;
  struct P { int K, mode, halo; double top, c; };
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is - p.halo, g.ie + p.halo, g.js - p.halo, g.je + p.halo)) return;
    if (p.mode == 0) {
      T s = T(p.top);
      x.out(0, 0, s);
#pragma unroll 4
      for (int k = 0; k < p.K; k++) { s = s + x.in(0, k); x.out(0, k + 1, s); }
    } else {
      T s = x.in(1, 0);
      x.out(0, p.K, s);
#pragma unroll 4
      for (int k = p.K - 1; k >= 0; k--) { s = s - p.c * x.in(0, k); x.out(0, k, s); }
    }
  }
  template <class X> DEV static void eval_ad(X& x, const P& p) {
    const Geom& g = x.g;
    if (!x.in_rect(g.is - p.halo, g.ie + p.halo, g.js - p.halo, g.je + p.halo)) return;
    double a = 0.0;
    // four levels per batch: their output adjoints and accumulators are requested together (distinct addresses), then updated
    if (p.mode == 0) {
      for (int k = p.K; k >= 1; k -= 4) {
        double o[4], t[4];
#pragma unroll
        for (int u = 0; u < 4; u++) if (k - u >= 1) { o[u] = x.oad(0, k - u); t[u] = x.iad(0, k - u - 1); }
#pragma unroll
        for (int u = 0; u < 4; u++) if (k - u >= 1) { a += o[u]; x.iad_set(0, k - u - 1, t[u] + a); }
      }
    } else {
      for (int k = 0; k < p.K; k += 4) {
        double o[4], t[4];
#pragma unroll
        for (int u = 0; u < 4; u++) if (k + u < p.K) { o[u] = x.oad(0, k + u); t[u] = x.iad(0, k + u); }
#pragma unroll
        for (int u = 0; u < 4; u++) if (k + u < p.K) { a += o[u]; x.iad_set(0, k + u, t[u] + (-a * p.c)); }
      }
      a += x.oad(0, p.K);
      x.add(1, 0, a);
    }
  }
};

// ---- column: tridiagonal solve --------------------------------------------------------------------------------
//   lo(k) x(k-1) + di(k) x(k) + up(k) x(k+1) = rhs(k),  k = 0..K-1
// in: lo di up rhs ; out: x.  lo / up are read at level k + lo_off / k + up_off (the w system passes the same
// interface array twice: lo = A(k), up = A(k+1)); lo_one: the sub-diagonal is 1.  The solution is written at level
// k + out_off (pp lives on interfaces 1..K; interface 0 is set to zero).
struct S_tri {
  static constexpr int NI = 4, NO = 1;
  struct P { int K, lo_one, lo_off, up_off, out_off, halo; };
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is - p.halo, g.ie + p.halo, g.js - p.halo, g.je + p.halo)) return;
    const int K = p.K;
    T gam[KMAX], xs[KMAX];
    T bet = x.in(1, 0);
    xs[0] = x.in(3, 0) / bet;
#pragma unroll 4
    for (int k = 1; k < K; k++) {
      T lo = p.lo_one ? T(1.0) : x.in(0, k + p.lo_off);
      gam[k] = x.in(2, k - 1 + p.up_off) / bet;
      bet = x.in(1, k) - lo * gam[k];
      xs[k] = (x.in(3, k) - lo * xs[k - 1]) / bet;
    }
    for (int k = K - 2; k >= 0; k--) xs[k] = xs[k] - gam[k + 1] * xs[k + 1];
    if (p.out_off) x.out(0, 0, T(0.0));
    for (int k = 0; k < K; k++) x.out(0, k + p.out_off, xs[k]);
  }
  // adjoint: A^T lam = x_ad ; rhs_ad = lam ; lo_ad(k) = -lam(k) x(k-1) ; di_ad(k) = -lam(k) x(k) ; up_ad(k) = -lam(k) x(k+1)
  // A^T has sub-diagonal up(k-1), diagonal di(k), super-diagonal lo(k+1)
  template <class X> DEV static void eval_ad(X& x, const P& p) {
    const Geom& g = x.g;
    if (!x.in_rect(g.is - p.halo, g.ie + p.halo, g.js - p.halo, g.je + p.halo)) return;
    const int K = p.K;
    double gam[KMAX], lam[KMAX];
    auto LO = [&](int k) { return p.lo_one ? 1.0 : x.in(0, k + p.lo_off); };
    auto UP = [&](int k) { return x.in(2, k + p.up_off); };
    double bet = x.in(1, 0);
    lam[0] = x.oad(0, p.out_off) / bet;
#pragma unroll 4
    for (int k = 1; k < K; k++) {
      gam[k] = LO(k) / bet;                       // super-diagonal of row k-1 of A^T is lo(k)
      bet = x.in(1, k) - UP(k - 1) * gam[k];      // sub-diagonal of row k is up(k-1)
      lam[k] = (x.oad(0, k + p.out_off) - UP(k - 1) * lam[k - 1]) / bet;
    }
    for (int k = K - 2; k >= 0; k--) lam[k] = lam[k] - gam[k + 1] * lam[k + 1];
    // per level: the four accumulators are requested together, then stored (lo and up may be the same interface array read at k + lo_off
    // and k + up_off: distinct addresses within a level as long as the offsets differ, else the two contributions are merged)
    const bool lo_up_same = x.same_ad(0, 2) && p.lo_off == p.up_off;
    // any other coincidence of accumulators (the same value bound to two inputs at the same level) keeps the plain sequence of add()s
    const bool alias = (!p.lo_one && (x.same_ad(0, 1) || x.same_ad(0, 3))) || x.same_ad(1, 2) || x.same_ad(1, 3) || x.same_ad(2, 3);
    if (alias) {
      for (int k = 0; k < K; k++) {
        const double l = lam[k], xk = x.outv(0, k + p.out_off);
        x.add(3, k, l);
        x.add(1, k, -l * xk);
        if (k >= 1 && !p.lo_one) x.add(0, k + p.lo_off, -l * x.outv(0, k - 1 + p.out_off));
        if (k < K - 1) x.add(2, k + p.up_off, -l * x.outv(0, k + 1 + p.out_off));
      }
      return;
    }
    for (int k = 0; k < K; k++) {
      const double l = lam[k], xk = x.outv(0, k + p.out_off);
      const bool has_lo = k >= 1 && !p.lo_one, has_up = k < K - 1;
      const double o3 = x.iad(3, k), o1 = x.iad(1, k);
      const double o0 = has_lo ? x.iad(0, k + p.lo_off) : 0.0, o2 = has_up ? x.iad(2, k + p.up_off) : 0.0;
      double d0 = has_lo ? -l * x.outv(0, k - 1 + p.out_off) : 0.0;
      const double d2 = has_up ? -l * x.outv(0, k + 1 + p.out_off) : 0.0;
      x.iad_set(3, k, o3 + l);
      x.iad_set(1, k, o1 + -l * xk);
      if (lo_up_same && has_lo && has_up) { x.iad_set(0, k + p.lo_off, o0 + d0 + d2); continue; }
      if (has_lo) x.iad_set(0, k + p.lo_off, o0 + d0);
      if (has_up) x.iad_set(2, k + p.up_off, o2 + d2);
    }
  }
};

// ---- column: pe2(0) = 0 ; pe2(k+1) = pe2(k) + dm(k) (w2(k) - w1(k)) / dt.   in: delp w2 w1 ; out: pe2 (K+1)
// SIM (SIM_solver, off-centred, model/nh_utils_nlm.F90:1434-1444): in: delp w2 w1 pp ;
//   pe2(k+1) = pe2(k) + (dm(k) (w2(k) - w1(k)) / dt - beta (pp(k+1) - pp(k))) ra
template <bool SIM> struct S_rs_pe2_t {
  static constexpr int NI = SIM ? 4 : 3, NO = 1;
  struct P { int K, halo; double rgrav, rdt, beta, ra; };
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is - p.halo, g.ie + p.halo, g.js - p.halo, g.je + p.halo)) return;
    T s = T(0.0);
    x.out(0, 0, s);
#pragma unroll 4
    for (int k = 0; k < p.K; k++) {
      if constexpr (SIM) s = s + ((x.in(0, k) * p.rgrav) * (x.in(1, k) - x.in(2, k)) * p.rdt - p.beta * (x.in(3, k + 1) - x.in(3, k))) * p.ra;
      else s = s + (x.in(0, k) * p.rgrav) * (x.in(1, k) - x.in(2, k)) * p.rdt;
      x.out(0, k + 1, s);
    }
  }
  template <class X> DEV static void eval_ad(X& x, const P& p) {
    const Geom& g = x.g;
    if (!x.in_rect(g.is - p.halo, g.ie + p.halo, g.js - p.halo, g.je + p.halo)) return;
    double a = 0.0;
    for (int k = p.K - 1; k >= 0; k--) {
      const double o0 = x.iad(0, k), o1 = x.iad(1, k), o2 = x.iad(2, k);     // requested together with the output adjoint, stored below
      a += x.oad(0, k + 1);
      const double dm = x.in(0, k) * p.rgrav, dw = x.in(1, k) - x.in(2, k);
      const double b = SIM ? a * p.ra : a;
      x.iad_set(0, k, o0 + b * p.rgrav * dw * p.rdt); x.iad_set(1, k, o1 + b * dm * p.rdt); x.iad_set(2, k, o2 + -b * dm * p.rdt);
      if constexpr (SIM) { x.add(3, k + 1, -b * p.beta); x.add(3, k, b * p.beta); }
    }
  }
};
using S_rs_pe2 = S_rs_pe2_t<false>;

// ---- column: p1(K-1) = (pe2(K-1) + 2 pe2(K))/3 ; p1(k) = (pe2(k) + bb(k) pe2(k+1) + g(k) pe2(k+2))/3 - g(k) p1(k+1)
// in: pe2 bb g ; out: p1 (K)
struct S_rs_p1 {
  static constexpr int NI = 3, NO = 1;
  struct P { int K, halo; };
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is - p.halo, g.ie + p.halo, g.js - p.halo, g.je + p.halo)) return;
    const int K = p.K;
    T p1 = (x.in(0, K - 1) + 2.0 * x.in(0, K)) * rmp::r3;
    x.out(0, K - 1, p1);
#pragma unroll 4
    for (int k = K - 2; k >= 0; k--) {
      T gk = x.in(2, k);
      p1 = (x.in(0, k) + x.in(1, k) * x.in(0, k + 1) + gk * x.in(0, k + 2)) * rmp::r3 - gk * p1;
      x.out(0, k, p1);
    }
  }
  template <class X> DEV static void eval_ad(X& x, const P& p) {
    const Geom& g = x.g;
    if (!x.in_rect(g.is - p.halo, g.ie + p.halo, g.js - p.halo, g.je + p.halo)) return;
    const int K = p.K;
    const double r3 = rmp::r3;
    double carry = 0.0;
    for (int k = 0; k < K - 1; k++) {
      // the five accumulators of the level (distinct addresses) are requested together, then stored
      const double a0 = x.iad(0, k), a01 = x.iad(0, k + 1), a02 = x.iad(0, k + 2), a1 = x.iad(1, k), a2 = x.iad(2, k);
      const double Pk = x.oad(0, k) + carry, gk = x.in(2, k);
      x.iad_set(0, k, a0 + Pk * r3); x.iad_set(1, k, a1 + Pk * r3 * x.in(0, k + 1)); x.iad_set(0, k + 1, a01 + Pk * r3 * x.in(1, k));
      x.iad_set(2, k, a2 + Pk * (r3 * x.in(0, k + 2) - x.outv(0, k + 1))); x.iad_set(0, k + 2, a02 + Pk * r3 * gk);
      carry = -gk * Pk;
    }
    const double Pk = x.oad(0, K - 1) + carry;
    x.add(0, K - 1, Pk * r3); x.add(0, K, 2.0 * Pk * r3);
  }
};

// ---- stencil: equation of state.  in: delp pt z pem ; out: pm2 pe           (nk = K; z, pem on K+1 interfaces)
struct S_rs_pe {
  static constexpr int NI = 4, NO = 2;
  struct P { int mode, halo; double gama, rgrav, rdgas; };
  static constexpr int NT = 6;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, 0, 0, 0}, {2, 0, 0, 1}, {3, 0, 0, 0}, {3, 0, 0, 1}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is - p.halo, g.ie + p.halo, g.js - p.halo, g.je + p.halo)) return;
    T dp = x.in(0), dz = x.in(2, 0, 0, 1) - x.in(2);
    T pm2 = (p.mode == 0) ? dp / m_log(x.in(3, 0, 0, 1) / x.in(3)) : dp / (m_log(x.in(3, 0, 0, 1)) - m_log(x.in(3)));
    x.out(0, pm2);
    x.out(1, m_exp(p.gama * m_log(-(dp * p.rgrav) / dz * p.rdgas * x.in(1))) - pm2);
  }
};

// ---- stencil: coefficients of the pp system.  in: delp pe ; out: bb g dd   (nk = K)
struct S_rs_cpp {
  static constexpr int NI = 2, NO = 3;
  struct P { int K, halo; };
  static constexpr int NT = 4;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {0, 0, 0, 1}, {1, 0, 0, 0}, {1, 0, 0, 1}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is - p.halo, g.ie + p.halo, g.js - p.halo, g.je + p.halo)) return;
    if (x.kk < p.K - 1) {
      T gk = x.in(0) / x.in(0, 0, 0, 1);
      x.out(0, 2.0 * (1.0 + gk)); x.out(1, gk); x.out(2, 3.0 * (x.in(1) + gk * x.in(1, 0, 0, 1)));
    } else {
      x.out(0, T(2.0)); x.out(1, T(0.0)); x.out(2, 3.0 * x.in(1));
    }
  }
};

// ---- stencil: off-diagonal of the w system on interfaces.  A(0) = 0 ; A(k) = t1g/(dz(k-1)+dz(k)) (pem(k)+pp(k)), 1 <= k < K ;
// A(K) = t1g/dz(K-1) (pem(K)+pp(K)).     in: z pem pp ; out: A   (nk = K+1)
struct S_rs_aa {
  static constexpr int NI = 3, NO = 1;
  struct P { int K, halo; double t1g; };
  static constexpr int NT = 5;
  static constexpr Tap taps[NT] = {{0, 0, 0, -1}, {0, 0, 0, 0}, {0, 0, 0, 1}, {1, 0, 0, 0}, {2, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is - p.halo, g.ie + p.halo, g.js - p.halo, g.je + p.halo)) return;
    const int k = x.kk;
    if (k == 0) { x.out(0, T(0.0)); return; }
    T dzm = x.in(0) - x.in(0, 0, 0, -1);                                   // dz(k-1)
    T den = (k < p.K) ? dzm + (x.in(0, 0, 0, 1) - x.in(0)) : dzm;          // dz(k-1) + dz(k) ; at k = K: dz(K-1)
    x.out(0, p.t1g / den * (x.in(1) + x.in(2)));
  }
};

// ---- stencil: diagonal and right-hand side of the w system.  in: delp w pp A ws ; out: di rhs   (nk = K)
// SIM (SIM_solver, model/nh_utils_nlm.F90:1396-1427): the explicit part of the off-centred scheme enters the right-hand side,
//   wk(k) = t2 A(k) (w(k-1) - w(k)) on interior interfaces, rhs(k) += wk(k+1) - wk(k) ; bottom layer: - wk(K-1) + A(K) (t2 w - ra ws)
template <bool SIM> struct S_rs_rw_t {
  static constexpr int NI = 5, NO = 2;
  struct P { int K, halo; double rgrav, dt, t2, ra; };
  static constexpr int NT = SIM ? 9 : 7;
  static constexpr Tap taps[9] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, 0, 0, 0}, {2, 0, 0, 1}, {3, 0, 0, 0}, {3, 0, 0, 1}, {4, 0, 0, KLAST},   // ws is 2-D: read by level K-1, owned by its single level
                                  {1, 0, 0, -1}, {1, 0, 0, 1}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is - p.halo, g.ie + p.halo, g.js - p.halo, g.je + p.halo)) return;
    T dm = x.in(0) * p.rgrav, a0 = x.in(3), a1 = x.in(3, 0, 0, 1);
    x.out(0, dm - (a0 + a1));
    T wc = x.in(1);
    T r = dm * wc + p.dt * (x.in(2, 0, 0, 1) - x.in(2));
    if constexpr (SIM) {
      if (x.kk < p.K - 1) r = r + p.t2 * a1 * (wc - x.in(1, 0, 0, 1));
      if (x.kk > 0) r = r - p.t2 * a0 * (x.in(1, 0, 0, -1) - wc);
      if (x.kk == p.K - 1) r = r + a1 * (p.t2 * wc - p.ra * x.in(4, 0, 0, KLAST));
    } else {
      if (x.kk == p.K - 1) r = r - a1 * x.in(4, 0, 0, KLAST);
    }
    x.out(1, r);
  }
};
using S_rs_rw = S_rs_rw_t<false>;

// ---- stencil: new layer thickness.  in: delp pt pm2 p1 ; out: dz   (nk = K)
struct S_rs_dz {
  static constexpr int NI = 4, NO = 1;
  struct P { int halo; double rgrav, rdgas, capa1, p_fac; };
  static constexpr int NT = 4;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, 0, 0, 0}, {3, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    const Geom& g = x.g;
    if (!x.in_rect(g.is - p.halo, g.ie + p.halo, g.js - p.halo, g.je + p.halo)) return;
    auto pm2 = x.in(2);
    x.out(0, -(x.in(0) * p.rgrav) * p.rdgas * x.in(1) * m_exp(p.capa1 * m_log(m_max(p.p_fac * pm2, x.in(3) + pm2))));
  }
};

// ---- stencil: full non-hydrostatic interface pressure of Riem_Solver_c.  in: pe2 pem ; out: pef   (nk = K+1)
struct S_rs_pef {
  static constexpr int NI = 2, NO = 1;
  struct P { int halo; double ptop; };
  static constexpr int NT = 2;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    const Geom& g = x.g;
    if (!x.in_rect(g.is - p.halo, g.ie + p.halo, g.js - p.halo, g.je + p.halo)) return;
    if (x.kk == 0) x.out(0, T(p.ptop)); else x.out(0, x.in(0) + x.in(1));
  }
};

// ---- stencil: SIM_solver's final blend  pe2 <- pe2 + beta (pp - pe2)  (model/nh_utils_nlm.F90:1460-1464).  in: pe2 pp ; out: ppe (nk = K+1)
struct S_rs_blend {
  static constexpr int NI = 2, NO = 1;
  struct P { int halo; double beta; };
  static constexpr int NT = 2;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    const Geom& g = x.g;
    if (!x.in_rect(g.is - p.halo, g.ie + p.halo, g.js - p.halo, g.je + p.halo)) return;
    auto pe2 = x.in(0);
    x.out(0, pe2 + p.beta * (x.in(1) - pe2));
  }
};

struct RiemOut { int pp, z, w, dz; };
// a_imp: Riem_Solver3 only (model/nh_core_nlm.F90:136-152): > 0.999 SIM1_solver, (0.5, 0.999] SIM_solver; Riem_Solver_c uses SIM1 for every a_imp > 0.5
struct RiemPrm { int K, mode, halo; double dt, akap, ptop, rdgas, grav, p_fac; double a_imp = 1.0; };
RiemOut build_riem(Program& P, const RiemPrm& r, int delp, int pt, int z, int w, int ws, int zb, const std::string& tag);

}  // namespace fv3lm
