// tp_core: 1-D PPM fluxes and the Lin-Rood 2-D flux-form transport operator fv_tp_2d.
// Reference: model_tlmadm/tp_core_tlm.F90  FV_TP_2D_TLM :2123, XPPM_TLM :2328, YPPM_TLM :2496,
// COPY_CORNERS_TLM :2843 (primal model/tp_core_nlm.F90:78-289, 291-420).  Only the linear
// orders the TL/AD implement are provided: iord = 1 (upwind), 2 (unlimited PPM) and 333 (third-order
// linear, Holdaway & Kent 2015) (tp_core_tlm.F90:2431-2488).
#pragma once
#include "engine.h"
#include "mosaic.h"

namespace fv3lm {

struct LevOrd { signed char v[128]; };   // per-level scheme order (sponge layers differ)
// hord = 333 (third-order linear scheme, tp_core_tlm.F90:2467-2488) is stored as ORD333 in a LevOrd
constexpr int ORD333 = 33;
// hord = 8 .. 13: the monotone PPM schemes of the nonlinear model (tp_core_nlm.F90:470-578), hord = 3 .. 7: its "smoothness-switch"
// schemes on the unlimited edge values (:327-467), stored as they are.  They are only legal on the trajectory side of a two-sided configuration: nothing is ever differentiated through them.
inline bool hord_is_mono(int hord) { return hord >= 3 && hord <= 13; }
inline int enc_hord(int hord, bool allow_mono = true) {
  if (hord == 1 || hord == 2) return hord;
  if (hord == 333) return ORD333;
  if (allow_mono && hord_is_mono(hord)) return hord;
  throw std::runtime_error("hord must be 1, 2 or 333 (the linear schemes the TL/AD implement, tp_core_tlm.F90:2431-2488)"
                           " or, for the trajectory of a two-sided configuration, 3..13 (tp_core_nlm.F90:327-578)");
}

namespace tp {
constexpr double p1 = 7.0 / 12.0, p2 = -1.0 / 12.0;
constexpr double c1 = -2.0 / 14.0, c2 = 11.0 / 14.0, c3 = 5.0 / 14.0;

// q at offset d along direction DIR (0 = x, 1 = y) of input f
template <int DIR, class X> DEV typename X::T Q(const X& x, int f, int d) {
  return DIR == 0 ? x.in(f, d, 0) : x.in(f, 0, d);
}
template <int DIR, class X> DEV double DA(const X& x, int d) {   // dxa / dya
  return DIR == 0 ? x.M(x.m.dxa, d, 0) : x.M(x.m.dya, 0, d);
}
// weight of q(e-2+n) in the edge value AL(e) on a cube edge (e = 1 or np), e at offset d from the current position: precomputed from
// dxa / dya when the metrics are uploaded (capi.cu), so that a warp with lanes on a cube edge loads four numbers instead of running two
// double-precision divisions per edge value (ncu r02k: the division paths were a quarter of the reverse tile kernel's instructions)
template <int DIR, class X> DEV double PW(const X& x, int n, int d) {
  const double* a = DIR == 0 ? (n == 0 ? x.m.ppmw_x0 : n == 1 ? x.m.ppmw_x1 : n == 2 ? x.m.ppmw_x2 : x.m.ppmw_x3)
                             : (n == 0 ? x.m.ppmw_y0 : n == 1 ? x.m.ppmw_y1 : n == 2 ? x.m.ppmw_y2 : x.m.ppmw_y3);
  return DIR == 0 ? x.M(a, d, 0) : x.M(a, 0, d);
}

// PPM edge value al at index (pos + d), pos = current i (DIR 0) or j (DIR 1).
// tp_core_tlm.F90:2396-2430 (x) and the symmetric y code.
template <int DIR, class X> DEV typename X::T edge_al(const X& x, int f, int d) {
  using T = typename X::T;
  const int ia = (DIR == 0 ? x.i : x.j) + d;
  const int np = DIR == 0 ? x.g.npx : x.g.npy;
  if (ia == 0 || ia == np - 1) return c1 * Q<DIR>(x, f, d - 2) + c2 * Q<DIR>(x, f, d - 1) + c3 * Q<DIR>(x, f, d);
  if (ia == 2 || ia == np + 1) return c3 * Q<DIR>(x, f, d - 1) + c2 * Q<DIR>(x, f, d) + c1 * Q<DIR>(x, f, d + 1);
  if (ia == 1 || ia == np)     // 0.5 (l + r) of the two one-sided extrapolations, as four precomputed weights
    return (PW<DIR>(x, 0, d) * Q<DIR>(x, f, d - 2) + PW<DIR>(x, 1, d) * Q<DIR>(x, f, d - 1)) + (PW<DIR>(x, 2, d) * Q<DIR>(x, f, d) + PW<DIR>(x, 3, d) * Q<DIR>(x, f, d + 1));
  return p1 * (Q<DIR>(x, f, d - 1) + Q<DIR>(x, f, d)) + p2 * (Q<DIR>(x, f, d - 2) + Q<DIR>(x, f, d + 1));
}

// the same edge value with the cube-edge extrapolation evaluated exactly as the reference writes it (two divisions): used by the
// smoothness-switch schemes of the nonlinear model (iord = 3 .. 7), whose discrete switches compare edge values and must see the reference's
// rounding on plateaus (tests/test_tp_core.py: integer-valued fields); the linear schemes have no switches and use the precomputed weights
template <int DIR, class X> DEV typename X::T edge_al_ref(const X& x, int f, int d) {
  using T = typename X::T;
  const int ia = (DIR == 0 ? x.i : x.j) + d;
  const int np = DIR == 0 ? x.g.npx : x.g.npy;
  if (ia == 1 || ia == np) {
    double a0 = DA<DIR>(x, d - 1), am = DA<DIR>(x, d - 2), a1 = DA<DIR>(x, d), a2 = DA<DIR>(x, d + 1);
    T l = ((2.0 * a0 + am) * Q<DIR>(x, f, d - 1) - a0 * Q<DIR>(x, f, d - 2)) / (am + a0);
    T r = ((2.0 * a1 + a2) * Q<DIR>(x, f, d) - a1 * Q<DIR>(x, f, d + 1)) / (a1 + a2);
    return 0.5 * (l + r);
  }
  return edge_al<DIR>(x, f, d);
}

// ---- monotone PPM (iord = 8 .. 13) in perturbation form: bl = AL - q, br = AR - q of the cell at offset co from the current face
// position.  tp_core_nlm.F90 xppm :470-569, yppm :780-893, pert_ppm :953-1012.  (dimensionless weights s11, s14, s15 :56)
constexpr double s11 = 11.0 / 14.0, s14 = 4.0 / 7.0, s15 = 3.0 / 14.0, r3 = 1.0 / 3.0, near_zero = 1.e-25, ppm_fac = 1.5;
template <class T> DEV T sgn_of(T a, T b) { T m = m_abs(a); return val(b) >= 0.0 ? m : T(0.0) - m; }      // Fortran SIGN(a, b)
template <class T> DEV T max3(T a, T b, T c) { return m_max(m_max(a, b), c); }
template <class T> DEV T min3(T a, T b, T c) { return m_min(m_min(a, b), c); }
// pert_ppm: iv = 0 positive-definite constraint, else the standard PPM monotonicity constraint
template <class T> DEV void pert_ppm(T a0, T& al, T& ar, int iv) {
  if (iv == 0) {
    if (val(a0) <= 0.0) { al = T(0.0); ar = T(0.0); return; }
    T a4 = -3.0 * (ar + al), da1 = ar - al;
    if (val(m_abs(da1)) < -val(a4)) {
      T fmin = a0 + 0.25 / a4 * (da1 * da1) + a4 * (1.0 / 12.0);
      if (val(fmin) < 0.0) {
        if (val(ar) > 0.0 && val(al) > 0.0) { ar = T(0.0); al = T(0.0); }
        else if (val(da1) > 0.0) ar = -2.0 * al;
        else al = -2.0 * ar;
      }
    }
  } else {
    if (val(al) * val(ar) < 0.0) {
      T da1 = al - ar, da2 = da1 * da1, a6da = 3.0 * (al + ar) * da1;
      if (val(a6da) < -val(da2)) ar = -2.0 * al;
      else if (val(a6da) > val(da2)) al = -2.0 * ar;
    } else { al = T(0.0); ar = T(0.0); }
  }
}
template <int DIR, class X> DEV void mono_blbr(const X& x, int fq, int co, int ord, typename X::T& bl, typename X::T& br) {
  using T = typename X::T;
  const int ia = (DIR == 0 ? x.i : x.j) + co, np = DIR == 0 ? x.g.npx : x.g.npy;
  auto q = [&](int d) { return Q<DIR>(x, fq, co + d); };                      // q of cell ia + d
  auto dm = [&](int d) {
    T qm = q(d - 1), q0 = q(d), qp = q(d + 1);
    T xt = 0.25 * (qp - qm);
    return sgn_of(min3(m_abs(xt), max3(qm, q0, qp) - q0, q0 - min3(qm, q0, qp)), xt);
  };
  auto al = [&](int d) { return 0.5 * (q(d - 1) + q(d)) + r3 * (dm(d - 1) - dm(d)); };   // west / south edge value of cell ia + d
  // mean of the two one-sided extrapolations to the cube edge between cells ia + d - 1 and ia + d, limited by the 4 cells around it
  auto edge = [&](int d) {
    const double am = DA<DIR>(x, co + d - 2), a0 = DA<DIR>(x, co + d - 1), a1 = DA<DIR>(x, co + d), a2 = DA<DIR>(x, co + d + 1);
    T xt = 0.5 * (((2.0 * a0 + am) * q(d - 1) - a0 * q(d - 2)) / (am + a0) + ((2.0 * a1 + a2) * q(d) - a1 * q(d + 1)) / (a1 + a2));
    xt = m_max(xt, m_min(m_min(q(d - 2), q(d - 1)), m_min(q(d), q(d + 1))));
    xt = m_min(xt, m_max(m_max(q(d - 2), q(d - 1)), m_max(q(d), q(d + 1))));
    return xt;
  };
  const T q0 = q(0);
  if (ia >= 3 && ia <= np - 3) {
    if (ord == 8 || ord == 11) {
      T xt = (ord == 8 ? 2.0 : ppm_fac) * dm(0);
      bl = T(0.0) - sgn_of(m_min(m_abs(xt), m_abs(al(0) - q0)), xt);
      br = sgn_of(m_min(m_abs(xt), m_abs(al(1) - q0)), xt);
    } else {
      bl = al(0) - q0; br = al(1) - q0;
      if (val(m_abs(dm(-1))) + val(m_abs(dm(0))) + val(m_abs(dm(1))) < near_zero) { bl = T(0.0); br = T(0.0); }
      else if (fabs(3.0 * (val(bl) + val(br))) > fabs(val(bl) - val(br))) {
        T pmp_2 = 2.0 * (q0 - q(-1)), lac_2 = pmp_2 - 0.75 * (2.0 * (q(-1) - q(-2)));
        br = m_min(max3(T(0.0), pmp_2, lac_2), m_max(br, min3(T(0.0), pmp_2, lac_2)));
        T pmp_1 = T(0.0) - 2.0 * (q(1) - q0), lac_1 = pmp_1 + 0.75 * (2.0 * (q(2) - q(1)));
        bl = m_min(max3(T(0.0), pmp_1, lac_1), m_max(bl, min3(T(0.0), pmp_1, lac_1)));
      }
    }
    if (ord == 9 || ord == 13) pert_ppm(q0, bl, br, 0);
    return;
  }
  if (ia == 0) { bl = s14 * dm(-1) + s11 * (q(-1) - q0); br = edge(1) - q0; }
  else if (ia == 1) { bl = edge(0) - q0; br = (s15 * q0 + s11 * q(1) - s14 * dm(1)) - q0; }
  else if (ia == 2) { bl = (s15 * q(-1) + s11 * q0 - s14 * dm(0)) - q0; br = al(1) - q0; }
  else if (ia == np - 2) { bl = al(0) - q0; br = (s15 * q(1) + s11 * q0 + s14 * dm(0)) - q0; }
  else if (ia == np - 1) { bl = (s15 * q0 + s11 * q(-1) + s14 * dm(-1)) - q0; br = edge(1) - q0; }
  else { bl = edge(0) - q0; br = s11 * (q(1) - q0) - s14 * dm(1); }     // ia == np
  pert_ppm(q0, bl, br, 1);
}

// 1-D flux at the current face from q (input fq) and Courant number c
// FULL = false: only the linear orders (1, 2, 333) -- the hot TL / AD kernels stay free of the nonlinear model's scheme code
// (which costs 10-20 registers per thread); FULL = true: all orders, used by the trajectory-only stages S_ppm_nl.
template <int DIR, bool FULL, class X> DEV typename X::T ppm_flux(const X& x, int fq, typename X::T c, int ord) {
  using T = typename X::T;
  if (ord == 1) return val(c) > 0.0 ? Q<DIR>(x, fq, -1) : Q<DIR>(x, fq, 0);
  if constexpr (FULL) {
  if (ord >= 8 && ord <= 13) {
    T bl, br;
    if (val(c) > 0.0) { mono_blbr<DIR>(x, fq, -1, ord, bl, br); return Q<DIR>(x, fq, -1) + (1.0 - c) * (br - c * (bl + br)); }
    mono_blbr<DIR>(x, fq, 0, ord, bl, br);
    return Q<DIR>(x, fq, 0) + (1.0 + c) * (bl + c * (bl + br));
  }
  if (ord >= 3 && ord <= 7) {
    // iord = 3 .. 7 (tp_core_nlm.F90:386-467): unlimited edge values, the second-order increment only where the profile is smooth
    auto AL = [&](int d) {
      T a = edge_al_ref<DIR>(x, fq, d);
      if (ord == 7 && val(a) < 0.0) {       // :343-346, :358-362, :369-373: positivity of the edge values
        const int ia = (DIR == 0 ? x.i : x.j) + d, np = DIR == 0 ? x.g.npx : x.g.npy;
        const bool edge = ia <= 2 || ia >= np - 1;
        a = edge ? T(0.0) : 0.5 * (Q<DIR>(x, fq, d - 1) + Q<DIR>(x, fq, d));
      }
      return a;
    };
    T qm = Q<DIR>(x, fq, -1), qp = Q<DIR>(x, fq, 0), a0 = AL(0);
    T blm = AL(-1) - qm, brm = a0 - qm, blp = a0 - qp, brp = AL(1) - qp;
    T b0m = blm + brm, b0p = blp + brp;
    bool sm5m, sm5p, sm6m = false, sm6p = false;
    if (ord == 3 || ord == 4) {
      sm5m = fabs(val(b0m)) < fabs(val(blm) - val(brm)); sm6m = 3.0 * fabs(val(b0m)) < fabs(val(blm) - val(brm));
      sm5p = fabs(val(b0p)) < fabs(val(blp) - val(brp)); sm6p = 3.0 * fabs(val(b0p)) < fabs(val(blp) - val(brp));
      if (ord == 3) {      // :386-414: falls back to the piece-wise linear increment before first order
        T fx1 = T(0.0);
        if (val(c) > 0.0) {
          if (sm6m || sm5p) fx1 = brm - c * b0m;
          else if (sm5m) fx1 = sgn_of(m_min(m_abs(blm), m_abs(brm)), brm);
          return qm + (1.0 - m_abs(c)) * fx1;
        }
        if (sm6p || sm5m) fx1 = blp + c * b0p;
        else if (sm5p) fx1 = sgn_of(m_min(m_abs(blp), m_abs(brp)), blp);
        return qp + (1.0 - m_abs(c)) * fx1;
      }
      if (val(c) > 0.0) return (sm6m || sm5p) ? qm + (1.0 - c) * (brm - c * b0m) : qm;
      return (sm6p || sm5m) ? qp + (1.0 + c) * (blp + c * b0p) : qp;
    }
    if (ord == 5) { sm5m = val(blm) * val(brm) < 0.0; sm5p = val(blp) * val(brp) < 0.0; }
    else { sm5m = fabs(3.0 * val(b0m)) < fabs(val(blm) - val(brm)); sm5p = fabs(3.0 * val(b0p)) < fabs(val(blp) - val(brp)); }
    (void)sm6m; (void)sm6p;
    if (val(c) > 0.0) return (sm5m || sm5p) ? qm + (1.0 - c) * (brm - c * b0m) : qm;
    return (sm5m || sm5p) ? qp + (1.0 + c) * (blp + c * b0p) : qp;
  }
  }   // FULL
  if (ord == ORD333) {
    // perfectly linear third-order scheme, no cube-edge special cases (tp_core_tlm.F90:2467-2488, :2638-2660)
    T qm1 = Q<DIR>(x, fq, -1), q0 = Q<DIR>(x, fq, 0);
    if (val(c) > 0.0) {
      T qm2 = Q<DIR>(x, fq, -2);
      return (2.0 * q0 + 5.0 * qm1 - qm2) / 6.0 - 0.5 * c * (q0 - qm1) + c * c / 6.0 * (q0 - 2.0 * qm1 + qm2);
    }
    T q1 = Q<DIR>(x, fq, 1);
    return (2.0 * qm1 + 5.0 * q0 - q1) / 6.0 - 0.5 * c * (q0 - qm1) + c * c / 6.0 * (q1 - 2.0 * q0 + qm1);
  }
  {
    // regular faces (all three edge values use the uniform weights): straight-line code, no edge tests
    const int ia = DIR == 0 ? x.i : x.j, np = DIR == 0 ? x.g.npx : x.g.npy;
    if (ia >= 4 && ia <= np - 3) {
      T qm2 = Q<DIR>(x, fq, -2), qm1 = Q<DIR>(x, fq, -1), q0 = Q<DIR>(x, fq, 0), q1 = Q<DIR>(x, fq, 1);
      T al0 = p1 * (qm1 + q0) + p2 * (qm2 + q1);
      if (val(c) > 0.0) {
        T alm = p1 * (qm2 + qm1) + p2 * (Q<DIR>(x, fq, -3) + q0);
        return qm1 + (1.0 - c) * (al0 - qm1 - c * (alm + al0 - (qm1 + qm1)));
      }
      T alp = p1 * (q0 + q1) + p2 * (qm1 + Q<DIR>(x, fq, 2));
      return q0 + (1.0 + c) * (al0 - q0 + c * (al0 + alp - (q0 + q0)));
    }
  }
  T al0 = edge_al<DIR>(x, fq, 0);
  if (val(c) > 0.0) {
    T qt = Q<DIR>(x, fq, -1);
    T alm = edge_al<DIR>(x, fq, -1);
    return qt + (1.0 - c) * (al0 - qt - c * (alm + al0 - (qt + qt)));
  } else {
    T qt = Q<DIR>(x, fq, 0);
    T alp = edge_al<DIR>(x, fq, 1);
    return qt + (1.0 + c) * (al0 - qt + c * (al0 + alp - (qt + qt)));
  }
}
}  // namespace tp

// flux(i,j) = xppm / yppm (q, c) on a rectangle.  in: 0 = q, 1 = c ; out: 0 = flux
template <int DIR> struct S_ppm {
  static constexpr int NI = 2, NO = 1;
  struct P { int i0, i1, j0, j1; LevOrd ord; };
  static constexpr int NT = 7;
  static constexpr Tap taps[NT] = {
      {0, DIR == 0 ? -3 : 0, DIR == 0 ? 0 : -3, 0}, {0, DIR == 0 ? -2 : 0, DIR == 0 ? 0 : -2, 0},
      {0, DIR == 0 ? -1 : 0, DIR == 0 ? 0 : -1, 0}, {0, 0, 0, 0},
      {0, DIR == 0 ? 1 : 0, DIR == 0 ? 0 : 1, 0},   {0, DIR == 0 ? 2 : 0, DIR == 0 ? 0 : 2, 0},
      {1, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    if (!x.in_rect(p.i0, p.i1, p.j0, p.j1)) return;
    x.out(0, tp::ppm_flux<DIR, false>(x, 0, x.in(1), p.ord.v[x.kk]));
  }

  // ---- hand-derived gather adjoint (replaces 7 seeded evaluations per cell).  The flux is linear in q:
  //   c > 0 : F = qt + (1-c)(al0 - qt - c(alm + al0 - 2 qt)),  qt = q(f-1), al0 = AL(f), alm = AL(f-1)
  //   c <= 0: F = qt + (1+c)(al0 - qt + c(al0 + alp - 2 qt)),  qt = q(f),   alp = AL(f+1)
  // with AL(e) = sum_n W_e[n] q(e-2+n) (tp::edge_al).  A cell p collects dF(f)/dq(p) * F_ad(f) from the faces
  // f = p-2 .. p+3, and the Courant number its own face's dF/dc * F_ad.
  static constexpr bool custom_ad = true;
  template <class X> DEV static double al_w(const X& x, int eo, int n) {   // weight of q(e-2+n) in AL(e), e = face + eo
    if (n < 0 || n > 3) return 0.0;
    const int ia = (DIR == 0 ? x.i : x.j) + eo;
    const int np = DIR == 0 ? x.g.npx : x.g.npy;
    if (ia == 0 || ia == np - 1) return n == 0 ? tp::c1 : n == 1 ? tp::c2 : n == 2 ? tp::c3 : 0.0;
    if (ia == 2 || ia == np + 1) return n == 0 ? 0.0 : n == 1 ? tp::c3 : n == 2 ? tp::c2 : tp::c1;
    if (ia == 1 || ia == np) return tp::PW<DIR>(x, n, eo);
    return (n == 0 || n == 3) ? tp::p2 : tp::p1;
  }
  template <class K> DEV static void adjoint(const K& kn, int ii, int jj, int kk, int tile, double* acc) {
    if (kk >= kn.nk_fwd || !kn.outad.p[0]) return;
    const P& p = kn.p;
    const int ord = p.ord.v[kk];
    CtxNL<S_ppm<DIR>> x; x.g = kn.g; x.m = kn.m; x.in_ = kn.in;
    const int nko = kn.outad.nk[0];
    x.setpos(ii, jj, kk, tile, kn.g.i0[tile], kn.g.j0[tile]);
    const int pos = DIR == 0 ? x.i : x.j, np = DIR == 0 ? kn.g.npx : kn.g.npy;
    const int loc = DIR == 0 ? x.il : x.jl, len = DIR == 0 ? kn.g.NX : kn.g.NY, arr = DIR == 0 ? ii : jj;
    // fast path: all six faces p-2..p+3 are regular (their three edge values use the uniform PPM weights), lie
    // inside the flux rectangle and inside the array -- true for every cell away from the cube edges
    const int f0 = DIR == 0 ? p.i0 : p.j0, f1 = DIR == 0 ? p.i1 : p.j1;
    const bool cross_ok = DIR == 0 ? (x.jl >= p.j0 && x.jl <= p.j1) : (x.il >= p.i0 && x.il <= p.i1);
    if (ord == 2 && cross_ok && pos - 3 >= 3 && pos + 4 <= np - 2 && loc - 2 >= f0 && loc + 3 <= f1 && arr - 2 >= 0 && arr + 3 < len) {
      const int stride = DIR == 0 ? 1 : kn.g.pitch;
      const int oc = x.off(kn.in.nk[1], 0, 0, 0), oa = x.off(nko, 0, 0, 0);
      const double* cp = kn.in.p[1]; const double* ap = kn.outad.p[0];
      double sum = 0.0;
#pragma unroll
      for (int s = -2; s <= 3; s++) {
        const double a = LDG(ap + oa + s * stride);
        const double c = LDG(cp + oc + s * stride);
        const int d = -s;
        // W = [p2, p1, p1, p2] at n = 0..3, zero outside
        auto W = [](int n) { return (n == 0 || n == 3) ? tp::p2 : ((n == 1 || n == 2) ? tp::p1 : 0.0); };
        double coef;
        if (c > 0.0) coef = (d == -1 ? 1.0 + (1.0 - c) * (2.0 * c - 1.0) : 0.0) + (1.0 - c) * (1.0 - c) * W(d + 2) - (1.0 - c) * c * W(d + 3);
        else coef = (d == 0 ? 1.0 - (1.0 + c) * (1.0 + 2.0 * c) : 0.0) + (1.0 + c) * (1.0 + c) * W(d + 2) + (1.0 + c) * c * W(d + 1);
        sum += coef * a;
      }
      acc[0] += sum;
    } else {
      for (int s = -2; s <= 3; s++) {           // face f = p + s reads this cell at offset d = -s
        const int fi = ii + (DIR == 0 ? s : 0), fj = jj + (DIR == 1 ? s : 0);
        if (fi < 0 || fi >= kn.g.NX || fj < 0 || fj >= kn.g.NY) continue;
        x.setpos(fi, fj, kk, tile, kn.g.i0[tile], kn.g.j0[tile]);
        if (!x.in_rect(p.i0, p.i1, p.j0, p.j1)) continue;
        const double a = kn.outad.p[0][x.off(nko, 0, 0, 0)];
        if (a == 0.0) continue;
        const double c = x.in(1);
        const int d = -s;
        double coef;
        if (ord == 1) coef = (c > 0.0) ? (d == -1 ? 1.0 : 0.0) : (d == 0 ? 1.0 : 0.0);
        else if (ord == ORD333) {
          const double c2 = c * c / 6.0;
          if (c > 0.0) coef = d == 0 ? 2.0 / 6.0 - 0.5 * c + c2 : d == -1 ? 5.0 / 6.0 + 0.5 * c - 2.0 * c2 : d == -2 ? -1.0 / 6.0 + c2 : 0.0;
          else coef = d == -1 ? 2.0 / 6.0 + 0.5 * c + c2 : d == 0 ? 5.0 / 6.0 - 0.5 * c - 2.0 * c2 : d == 1 ? -1.0 / 6.0 + c2 : 0.0;
        } else if (c > 0.0) {
          const double dqt = 1.0 + (1.0 - c) * (2.0 * c - 1.0), dal0 = (1.0 - c) * (1.0 - c), dalm = -(1.0 - c) * c;
          coef = (d == -1 ? dqt : 0.0) + dal0 * al_w(x, 0, d + 2) + dalm * al_w(x, -1, d + 3);
        } else {
          const double dqt = 1.0 - (1.0 + c) * (1.0 + 2.0 * c), dal0 = (1.0 + c) * (1.0 + c), dalp = (1.0 + c) * c;
          coef = (d == 0 ? dqt : 0.0) + dal0 * al_w(x, 0, d + 2) + dalp * al_w(x, 1, d + 1);
        }
        acc[0] += coef * a;
      }
    }
    // Courant number: the face at this cell
    if (kn.inad.p[1] && ord != 1) {
      x.setpos(ii, jj, kk, tile, kn.g.i0[tile], kn.g.j0[tile]);
      if (x.in_rect(p.i0, p.i1, p.j0, p.j1)) {
        const double a = kn.outad.p[0][x.off(nko, 0, 0, 0)];
        if (a != 0.0) {
          const double c = x.in(1);
          if (ord == ORD333) {
            const double qm1 = tp::Q<DIR>(x, 0, -1), q0 = tp::Q<DIR>(x, 0, 0);
            const double curv = c > 0.0 ? q0 - 2.0 * qm1 + tp::Q<DIR>(x, 0, -2) : tp::Q<DIR>(x, 0, 1) - 2.0 * q0 + qm1;
            acc[1] += (-0.5 * (q0 - qm1) + c / 3.0 * curv) * a;
            return;
          }
          double al0, qt, al2;      // al2: the upwind-side second edge value (alm for c > 0, alp otherwise)
          if (pos >= 4 && pos <= np - 3) {
            const double qm2 = tp::Q<DIR>(x, 0, -2), qm1 = tp::Q<DIR>(x, 0, -1), q0 = tp::Q<DIR>(x, 0, 0), q1 = tp::Q<DIR>(x, 0, 1);
            al0 = tp::p1 * (qm1 + q0) + tp::p2 * (qm2 + q1);
            if (c > 0.0) { qt = qm1; al2 = tp::p1 * (qm2 + qm1) + tp::p2 * (tp::Q<DIR>(x, 0, -3) + q0); }
            else { qt = q0; al2 = tp::p1 * (q0 + q1) + tp::p2 * (qm1 + tp::Q<DIR>(x, 0, 2)); }
          } else {
            al0 = tp::edge_al<DIR>(x, 0, 0);
            if (c > 0.0) { qt = tp::Q<DIR>(x, 0, -1); al2 = tp::edge_al<DIR>(x, 0, -1); }
            else { qt = tp::Q<DIR>(x, 0, 0); al2 = tp::edge_al<DIR>(x, 0, 1); }
          }
          double dc;
          if (c > 0.0) { const double b = al2 + al0 - (qt + qt); dc = -(al0 - qt - c * b) - (1.0 - c) * b; }
          else { const double b = al0 + al2 - (qt + qt); dc = (al0 - qt + c * b) + (1.0 + c) * b; }
          acc[1] += dc * a;
        }
      }
    }
  }
};

// the same flux with every order of the nonlinear model (hord 1 .. 13, 333): trajectory side of a two-sided configuration.
// Its inputs are detached views, so only the value-only kernel ever runs; the taps merely describe the footprint.
template <int DIR> struct S_ppm_nl {
  static constexpr int NI = 2, NO = 1;
  using P = typename S_ppm<DIR>::P;
  static constexpr int NT = 7;
  static constexpr Tap taps[NT] = {
      {0, DIR == 0 ? -3 : 0, DIR == 0 ? 0 : -3, 0}, {0, DIR == 0 ? -2 : 0, DIR == 0 ? 0 : -2, 0},
      {0, DIR == 0 ? -1 : 0, DIR == 0 ? 0 : -1, 0}, {0, 0, 0, 0},
      {0, DIR == 0 ? 1 : 0, DIR == 0 ? 0 : 1, 0},   {0, DIR == 0 ? 2 : 0, DIR == 0 ? 0 : 2, 0},
      {1, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    if (!x.in_rect(p.i0, p.i1, p.j0, p.j1)) return;
    x.out(0, tp::ppm_flux<DIR, true>(x, 0, x.in(1), p.ord.v[x.kk]));
  }
};
inline bool ord_is_linear(const LevOrd& o, int nk) {
  for (int k = 0; k < nk; k++) if (o.v[k] != 1 && o.v[k] != 2 && o.v[k] != ORD333) return false;
  return true;
}

// inner update  q_i = (q*area + fyy(j) - fyy(j+1)) / ra_y ,  fyy = yfx * fy2   (DIR = 1)
//               q_j = (q*area + fx1(i) - fx1(i+1)) / ra_x ,  fx1 = xfx * fx2   (DIR = 0)
// in: 0 = q, 1 = inner flux (fy2/fx2), 2 = yfx/xfx, 3 = ra ; out: 0 = q_i/q_j
template <int DIR> struct S_inner {
  static constexpr int NI = 4, NO = 1;
  struct P { int i0, i1, j0, j1; };
  static constexpr int NT = 6;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {1, DIR == 0, DIR == 1, 0},
                                   {2, 0, 0, 0}, {2, DIR == 0, DIR == 1, 0}, {3, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    if (!x.in_rect(p.i0, p.i1, p.j0, p.j1)) return;
    using T = typename X::T;
    constexpr int di = DIR == 0, dj = DIR == 1;
    T f0 = x.in(2) * x.in(1), f1 = x.in(2, di, dj) * x.in(1, di, dj);
    x.out(0, (x.in(0) * x.M(x.m.area) + f0 - f1) / x.in(3));
  }
};

// flux averaging  f = 0.5*(f_outer + f_inner) * m   (tp_core_tlm.F90:2268-2313)
// in: 0 = outer flux, 1 = inner flux, 2 = multiplier (xfx/yfx or mfx/mfy) ; out: 0
struct S_favg {
  static constexpr int NI = 3, NO = 1;
  struct P { int i0, i1, j0, j1; };
  static constexpr int NT = 3;
  static constexpr Tap taps[NT] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, 0, 0, 0}};
  template <class X> DEV static void eval(X& x, const P& p) {
    if (!x.in_rect(p.i0, p.i1, p.j0, p.j1)) return;
    x.out(0, 0.5 * (x.in(0) + x.in(1)) * x.in(2));
  }
};

// Builder: appends fv_tp_2d to a program.  q is patched in place (copy_corners).
// mfx/mfy < 0 selects the "delp / vorticity" form (multiplier = xfx/yfx).
struct TpOut { int fx, fy; };
TpOut build_fv_tp_2d(Program& P, Mosaic& mo, int q, int crx, int cry, int xfx, int yfx, int ra_x, int ra_y,
                     int mfx, int mfy, const LevOrd& hord, int nk, const std::string& tag);

}  // namespace fv3lm
