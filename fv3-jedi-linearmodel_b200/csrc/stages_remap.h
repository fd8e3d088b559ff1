// Vertical remap (model/fv_mapz_nlm.F90): cs_profile / scalar_profile for |kord| > 16
// (:2113-2200, :1730-1813) + the overlap integration of map_scalar / map1_ppm / map1_q2
// (:1237-1330, :1332-1424, :1541-1634); tracer_2d helper stages (fv_tracer2d_nlm.F90:275-516).
// TL: model_tlmadm/fv_mapz_tlm.F90 MAP_SCALAR_TLM :7765, MAP1_PPM_TLM :7909, MAP1_Q2_TLM :8222,
// CS_PROFILE_TLM :8513; AD fv_mapz_adm.F90.  The adjoint is a hand-written column reverse sweep.
#pragma once
#include "engine.h"
#include "mosaic.h"

namespace fv3lm {

constexpr int KMAX = 96;

namespace rmp {
constexpr double r3 = 1.0 / 3.0, r23 = 2.0 / 3.0;

// interface values q[0..K] (q[k] = top edge of layer k, 0-based) of the cubic-spline profile
template <class T> DEV void cs_profile(const T* a, const T* dp, T* q, T* gam, int K, int iv, T qs) {
  if (iv == -2) {
    // gam is shifted by one in the reference (gam(k+1)); store gam[k] = reference gam(k+1), k 0-based
    gam[1] = T(0.5);
    q[0] = 1.5 * a[0];
    for (int k = 1; k < K - 1; k++) {
      T grat = dp[k - 1] / dp[k];
      T bet = 2.0 + grat + grat - gam[k];
      q[k] = (3.0 * (a[k - 1] + a[k]) - q[k - 1]) / bet;
      gam[k + 1] = grat / bet;
    }
    T grat = dp[K - 2] / dp[K - 1];
    q[K - 1] = (3.0 * (a[K - 2] + a[K - 1]) - grat * qs - q[K - 2]) / (2.0 + grat + grat - gam[K - 1]);
    q[K] = qs;
    for (int k = K - 2; k >= 0; k--) q[k] = q[k] - gam[k + 1] * q[k + 1];
  } else {
    T grat = dp[1] / dp[0];
    T bet = grat * (grat + 0.5);
    q[0] = ((grat + grat) * (grat + 1.0) * a[0] + a[1]) / bet;
    gam[0] = (1.0 + grat * (grat + 1.5)) / bet;
    T d4 = T(0.0);
    for (int k = 1; k < K; k++) {
      d4 = dp[k - 1] / dp[k];
      bet = 2.0 + d4 + d4 - gam[k - 1];
      q[k] = (3.0 * (a[k - 1] + d4 * a[k]) - q[k - 1]) / bet;
      gam[k] = d4 / bet;
    }
    T a_bot = 1.0 + d4 * (d4 + 1.5);
    q[K] = (2.0 * d4 * (d4 + 1.0) * a[K - 1] + a[K - 2] - a_bot * q[K - 1]) / (d4 * (d4 + 0.5) - a_bot * gam[K - 1]);
    for (int k = K - 1; k >= 0; k--) q[k] = q[k] - gam[k] * q[k + 1];
  }
}

// layer (0-based) of source grid pe1 containing p, reference tie rule: first l with pe1[l] <= p <= pe1[l+1]
DEV int find_layer(const double* pe1v, double p, int l0, int K) {
  int l = l0;
  while (l < K - 1 && !(p >= pe1v[l] && p <= pe1v[l + 1])) l++;
  return l;
}
}  // namespace rmp

// Remap one field.   in: a pe1 pe2 qs dp2 ; out: q2
//   iv = -2 : bottom boundary value qs (2-D field) is used (w);  use_dp2: divide by dp2 (map1_q2)
struct S_remap {
  static constexpr int NI = 5, NO = 1;
  struct P { int K, iv, use_dp2; int i0, i1, j0, j1; };
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    if (!x.in_rect(p.i0, p.i1, p.j0, p.j1)) return;
    const int K = p.K;
    T a[KMAX], pe1[KMAX + 1], pe2[KMAX + 1], dp[KMAX], q[KMAX + 1], gam[KMAX + 1];
    double pe1v[KMAX + 1];
    for (int k = 0; k <= K; k++) { pe1[k] = x.in(1, k); pe2[k] = x.in(2, k); pe1v[k] = val(pe1[k]); }
    for (int k = 0; k < K; k++) { a[k] = x.in(0, k); dp[k] = pe1[k + 1] - pe1[k]; }
    T qs = (p.iv == -2) ? x.in(3, 0) : T(0.0);
    rmp::cs_profile(a, dp, q, gam, K, p.iv, qs);
    int k0 = 0;
    for (int k = 0; k < K; k++) {
      int l = rmp::find_layer(pe1v, val(pe2[k]), k0, K);
      T a2 = q[l], a3 = q[l + 1], a4 = 3.0 * (2.0 * a[l] - (a2 + a3));
      T pl = (pe2[k] - pe1[l]) / dp[l];
      T res;
      if (val(pe2[k + 1]) <= pe1v[l + 1]) {
        T pr = (pe2[k + 1] - pe1[l]) / dp[l];
        res = a2 + 0.5 * (a4 + a3 - a2) * (pr + pl) - a4 * rmp::r3 * (pr * (pr + pl) + pl * pl);
        k0 = l;
      } else {
        T qsum = (pe1[l + 1] - pe2[k]) * (a2 + 0.5 * (a4 + a3 - a2) * (1.0 + pl) - a4 * (rmp::r3 * (1.0 + pl * (1.0 + pl))));
        for (int m = l + 1; m < K; m++) {
          if (val(pe2[k + 1]) > pe1v[m + 1]) {
            qsum = qsum + dp[m] * a[m];
          } else {
            T dpl = pe2[k + 1] - pe1[m];
            T esl = dpl / dp[m];
            T b2 = q[m], b3 = q[m + 1], b4 = 3.0 * (2.0 * a[m] - (b2 + b3));
            qsum = qsum + dpl * (b2 + 0.5 * esl * (b3 - b2 + b4 * (1.0 - rmp::r23 * esl)));
            k0 = m;
            break;
          }
        }
        T den = p.use_dp2 ? x.in(4, k) : pe2[k + 1] - pe2[k];
        res = qsum / den;
      }
      x.out(0, k, res);
    }
  }

  // ---- hand-written adjoint: recompute the forward column in double, then reverse
  template <class X> DEV static void eval_ad(X& x, const P& p) {
    if (!x.in_rect(p.i0, p.i1, p.j0, p.j1)) return;
    const int K = p.K;
    double a[KMAX], pe1[KMAX + 1], pe2[KMAX + 1], dp[KMAX], q[KMAX + 1], gam[KMAX + 1];
    double qt[KMAX + 1], bet[KMAX], d4[KMAX];
    double a_ad[KMAX], q_ad[KMAX + 1], pe1_ad[KMAX + 1], pe2_ad[KMAX + 1], dp_ad[KMAX], gam_ad[KMAX + 1], dp2_ad[KMAX];
    for (int k = 0; k <= K; k++) { pe1[k] = x.in(1, k); pe2[k] = x.in(2, k); q_ad[k] = 0.0; pe1_ad[k] = 0.0; pe2_ad[k] = 0.0; gam_ad[k] = 0.0; }
    for (int k = 0; k < K; k++) { a[k] = x.in(0, k); dp[k] = pe1[k + 1] - pe1[k]; a_ad[k] = 0.0; dp_ad[k] = 0.0; dp2_ad[k] = 0.0; }
    const double qs = (p.iv == -2) ? x.in(3, 0) : 0.0;
    double qs_ad = 0.0;
    // forward profile keeping the elimination intermediates
    if (p.iv == -2) {
      gam[1] = 0.5; qt[0] = 1.5 * a[0];
      for (int k = 1; k < K - 1; k++) {
        d4[k] = dp[k - 1] / dp[k];
        bet[k] = 2.0 + d4[k] + d4[k] - gam[k];
        qt[k] = (3.0 * (a[k - 1] + a[k]) - qt[k - 1]) / bet[k];
        gam[k + 1] = d4[k] / bet[k];
      }
      d4[K - 1] = dp[K - 2] / dp[K - 1];
      bet[K - 1] = 2.0 + d4[K - 1] + d4[K - 1] - gam[K - 1];
      qt[K - 1] = (3.0 * (a[K - 2] + a[K - 1]) - d4[K - 1] * qs - qt[K - 2]) / bet[K - 1];
      q[K] = qs; q[K - 1] = qt[K - 1];
      for (int k = K - 2; k >= 0; k--) q[k] = qt[k] - gam[k + 1] * q[k + 1];
    } else {
      double grat = dp[1] / dp[0];
      bet[0] = grat * (grat + 0.5);
      qt[0] = ((grat + grat) * (grat + 1.0) * a[0] + a[1]) / bet[0];
      gam[0] = (1.0 + grat * (grat + 1.5)) / bet[0];
      d4[0] = grat;
      for (int k = 1; k < K; k++) {
        d4[k] = dp[k - 1] / dp[k];
        bet[k] = 2.0 + d4[k] + d4[k] - gam[k - 1];
        qt[k] = (3.0 * (a[k - 1] + d4[k] * a[k]) - qt[k - 1]) / bet[k];
        gam[k] = d4[k] / bet[k];
      }
      double d = d4[K - 1], a_bot = 1.0 + d * (d + 1.5);
      q[K] = (2.0 * d * (d + 1.0) * a[K - 1] + a[K - 2] - a_bot * qt[K - 1]) / (d * (d + 0.5) - a_bot * gam[K - 1]);
      for (int k = K - 1; k >= 0; k--) q[k] = qt[k] - gam[k] * q[k + 1];
    }
    // ---- reverse of the integration (target layers are independent)
    auto a4ad = [&](int l, double v) { a_ad[l] += 6.0 * v; q_ad[l] -= 3.0 * v; q_ad[l + 1] -= 3.0 * v; };
    int k0 = 0;
    for (int k = 0; k < K; k++) {
      const double r_ad = x.oad(0, k);
      int l = rmp::find_layer(pe1, pe2[k], k0, K);
      double a2 = q[l], a3 = q[l + 1], a4 = 3.0 * (2.0 * a[l] - (a2 + a3));
      double pl = (pe2[k] - pe1[l]) / dp[l];
      double pl_ad = 0.0;
      if (pe2[k + 1] <= pe1[l + 1]) {
        double pr = (pe2[k + 1] - pe1[l]) / dp[l], s = pr + pl;
        double c = 0.5 * (a4 + a3 - a2);
        q_ad[l] += r_ad * (1.0 - 0.5 * s); q_ad[l + 1] += r_ad * 0.5 * s;
        a4ad(l, r_ad * (0.5 * s - rmp::r3 * (pr * s + pl * pl)));
        double pr_ad = r_ad * (c - a4 * rmp::r3 * (2.0 * pr + pl));
        pl_ad = r_ad * (c - a4 * rmp::r3 * (pr + 2.0 * pl));
        pe2_ad[k + 1] += pr_ad / dp[l]; pe1_ad[l] -= pr_ad / dp[l]; dp_ad[l] -= pr_ad * pr / dp[l];
        k0 = l;
      } else {
        // recompute qsum pieces
        double F = a2 + 0.5 * (a4 + a3 - a2) * (1.0 + pl) - a4 * (rmp::r3 * (1.0 + pl * (1.0 + pl)));
        double qsum = (pe1[l + 1] - pe2[k]) * F;
        int mlast = -1;
        for (int m = l + 1; m < K; m++) {
          if (pe2[k + 1] > pe1[m + 1]) qsum += dp[m] * a[m];
          else {
            double dpl = pe2[k + 1] - pe1[m], esl = dpl / dp[m];
            double b2 = q[m], b3 = q[m + 1], b4 = 3.0 * (2.0 * a[m] - (b2 + b3));
            qsum += dpl * (b2 + 0.5 * esl * (b3 - b2 + b4 * (1.0 - rmp::r23 * esl)));
            mlast = m;
            break;
          }
        }
        double den = p.use_dp2 ? x.in(4, k) : pe2[k + 1] - pe2[k];
        double qs_ad_ = r_ad / den;                 // adjoint of qsum
        double den_ad = -r_ad * (qsum / den) / den;
        if (p.use_dp2) dp2_ad[k] += den_ad; else { pe2_ad[k + 1] += den_ad; pe2_ad[k] -= den_ad; }
        // first piece
        double w = pe1[l + 1] - pe2[k];
        pe1_ad[l + 1] += qs_ad_ * F; pe2_ad[k] -= qs_ad_ * F;
        double F_ad = qs_ad_ * w;
        q_ad[l] += F_ad * (1.0 - 0.5 * (1.0 + pl)); q_ad[l + 1] += F_ad * 0.5 * (1.0 + pl);
        a4ad(l, F_ad * (0.5 * (1.0 + pl) - rmp::r3 * (1.0 + pl + pl * pl)));
        pl_ad = F_ad * (0.5 * (a4 + a3 - a2) - a4 * rmp::r3 * (1.0 + 2.0 * pl));
        // full layers and the last partial layer
        const int mend = (mlast >= 0) ? mlast : K;
        for (int m = l + 1; m < mend; m++) { dp_ad[m] += qs_ad_ * a[m]; a_ad[m] += qs_ad_ * dp[m]; }
        if (mlast >= 0) {
          int m = mlast;
          double dpl = pe2[k + 1] - pe1[m], esl = dpl / dp[m];
          double b2 = q[m], b3 = q[m + 1], b4 = 3.0 * (2.0 * a[m] - (b2 + b3));
          double G = b2 + 0.5 * esl * (b3 - b2 + b4 * (1.0 - rmp::r23 * esl));
          double dpl_ad = qs_ad_ * G, G_ad = qs_ad_ * dpl;
          q_ad[m] += G_ad * (1.0 - 0.5 * esl); q_ad[m + 1] += G_ad * 0.5 * esl;
          a4ad(m, G_ad * 0.5 * esl * (1.0 - rmp::r23 * esl));
          double esl_ad = G_ad * 0.5 * (b3 - b2 + b4 * (1.0 - 2.0 * rmp::r23 * esl));
          dpl_ad += esl_ad / dp[m]; dp_ad[m] -= esl_ad * esl / dp[m];
          pe2_ad[k + 1] += dpl_ad; pe1_ad[m] -= dpl_ad;
          k0 = m;
        }
      }
      // pl = (pe2[k] - pe1[l]) / dp[l]
      pe2_ad[k] += pl_ad / dp[l]; pe1_ad[l] -= pl_ad / dp[l]; dp_ad[l] -= pl_ad * pl / dp[l];
    }
    // ---- reverse of the profile
    double qt_ad[KMAX + 1];
    for (int k = 0; k <= K; k++) qt_ad[k] = 0.0;
    if (p.iv == -2) {
      for (int k = 0; k <= K - 2; k++) {            // q[k] = qt[k] - gam[k+1] q[k+1]
        qt_ad[k] += q_ad[k]; gam_ad[k + 1] -= q_ad[k] * q[k + 1]; q_ad[k + 1] -= gam[k + 1] * q_ad[k];
      }
      qt_ad[K - 1] += q_ad[K - 1]; qs_ad += q_ad[K];
      {  // qt[K-1] = (3(a[K-2]+a[K-1]) - d4 qs - qt[K-2]) / bet
        int k = K - 1;
        double n_ad = qt_ad[k] / bet[k], bet_ad = -qt_ad[k] * qt[k] / bet[k];
        a_ad[k - 1] += 3.0 * n_ad; a_ad[k] += 3.0 * n_ad; qs_ad -= n_ad * d4[k]; qt_ad[k - 1] -= n_ad;
        double d4_ad = -n_ad * qs + 2.0 * bet_ad;
        gam_ad[k] -= bet_ad;
        dp_ad[k - 1] += d4_ad / dp[k]; dp_ad[k] -= d4_ad * d4[k] / dp[k];
      }
      for (int k = K - 2; k >= 1; k--) {
        double d4_ad = gam_ad[k + 1] / bet[k], bet_ad = -gam_ad[k + 1] * gam[k + 1] / bet[k];
        double n_ad = qt_ad[k] / bet[k]; bet_ad -= qt_ad[k] * qt[k] / bet[k];
        a_ad[k - 1] += 3.0 * n_ad; a_ad[k] += 3.0 * n_ad; qt_ad[k - 1] -= n_ad;
        d4_ad += 2.0 * bet_ad; gam_ad[k] -= bet_ad;
        dp_ad[k - 1] += d4_ad / dp[k]; dp_ad[k] -= d4_ad * d4[k] / dp[k];
      }
      a_ad[0] += 1.5 * qt_ad[0];
      x.add(3, 0, qs_ad);
    } else {
      for (int k = 0; k < K; k++) {                 // q[k] = qt[k] - gam[k] q[k+1]
        qt_ad[k] += q_ad[k]; gam_ad[k] -= q_ad[k] * q[k + 1]; q_ad[k + 1] -= gam[k] * q_ad[k];
      }
      double d = d4[K - 1], a_bot = 1.0 + d * (d + 1.5), den = d * (d + 0.5) - a_bot * gam[K - 1];
      double num_ad = q_ad[K] / den, den_ad = -q_ad[K] * q[K] / den;
      double d_ad = num_ad * 2.0 * (2.0 * d + 1.0) * a[K - 1] + den_ad * (2.0 * d + 0.5);
      a_ad[K - 1] += num_ad * 2.0 * d * (d + 1.0); a_ad[K - 2] += num_ad;
      double abot_ad = -num_ad * qt[K - 1] - den_ad * gam[K - 1];
      qt_ad[K - 1] -= num_ad * a_bot; gam_ad[K - 1] -= den_ad * a_bot;
      d_ad += abot_ad * (2.0 * d + 1.5);
      double d4_carry = d_ad;                       // adjoint of d4[K-1] from the bottom closure
      for (int k = K - 1; k >= 1; k--) {
        double d4_ad = (k == K - 1 ? d4_carry : 0.0) + gam_ad[k] / bet[k];
        double bet_ad = -gam_ad[k] * gam[k] / bet[k];
        double n_ad = qt_ad[k] / bet[k]; bet_ad -= qt_ad[k] * qt[k] / bet[k];
        a_ad[k - 1] += 3.0 * n_ad; d4_ad += 3.0 * n_ad * a[k]; a_ad[k] += 3.0 * n_ad * d4[k]; qt_ad[k - 1] -= n_ad;
        d4_ad += 2.0 * bet_ad; gam_ad[k - 1] -= bet_ad;
        dp_ad[k - 1] += d4_ad / dp[k]; dp_ad[k] -= d4_ad * d4[k] / dp[k];
      }
      double grat = d4[0];
      double g_ad = gam_ad[0];
      double grat_ad = g_ad * (2.0 * grat + 1.5) / bet[0], b_ad = -g_ad * gam[0] / bet[0];
      double n_ad = qt_ad[0] / bet[0]; b_ad -= qt_ad[0] * qt[0] / bet[0];
      grat_ad += n_ad * (4.0 * grat + 2.0) * a[0]; a_ad[0] += n_ad * 2.0 * grat * (grat + 1.0); a_ad[1] += n_ad;
      grat_ad += b_ad * (2.0 * grat + 0.5);
      dp_ad[1] += grat_ad / dp[0]; dp_ad[0] -= grat_ad * grat / dp[0];
    }
    for (int k = 0; k < K; k++) { pe1_ad[k + 1] += dp_ad[k]; pe1_ad[k] -= dp_ad[k]; }
    // flush: four levels per batch, accumulators requested before the stores (see ColAD::iad)
    for (int k = 0; k < K; k += 4) {
      double t0[4], t4[4];
#pragma unroll
      for (int u = 0; u < 4; u++) if (k + u < K) { t0[u] = x.iad(0, k + u); t4[u] = p.use_dp2 ? x.iad(4, k + u) : 0.0; }
#pragma unroll
      for (int u = 0; u < 4; u++) if (k + u < K) { x.iad_set(0, k + u, t0[u] + a_ad[k + u]); if (p.use_dp2) x.iad_set(4, k + u, t4[u] + dp2_ad[k + u]); }
    }
    if (x.same_ad(1, 2)) {
      for (int k = 0; k <= K; k++) { x.add(1, k, pe1_ad[k]); x.add(2, k, pe2_ad[k]); }
    } else {
      for (int k = 0; k <= K; k += 4) {
        double t1[4], t2[4];
#pragma unroll
        for (int u = 0; u < 4; u++) if (k + u <= K) { t1[u] = x.iad(1, k + u); t2[u] = x.iad(2, k + u); }
#pragma unroll
        for (int u = 0; u < 4; u++) if (k + u <= K) { x.iad_set(1, k + u, t1[u] + pe1_ad[k + u]); x.iad_set(2, k + u, t2[u] + pe2_ad[k + u]); }
      }
    }
  }
};

// ---------------------------------------------------------------------------------
// Monotone remap of the nonlinear model (|kord| = 8 .. 14): the limited sub-grid profile of scalar_profile / cs_profile
// (model/fv_mapz_nlm.F90:1814-2109, :2197-2463) and cs_limiters (:2467-2542).  Trajectory side of a two-sided configuration only
// (split_kord, model_tlmadm/fv_arrays_tlmadm.F90:69): its inputs are detached views, nothing is differentiated through it.
// ---------------------------------------------------------------------------------
namespace rmp {
template <class T> DEV T mn3(T a, T b, T c) { return m_min(m_min(a, b), c); }
template <class T> DEV T mx3(T a, T b, T c) { return m_max(m_max(a, b), c); }
template <class T> DEV void cs_limiters(bool extm, T a, T& a2, T& a3, T& a4, int iv) {
  if (iv == 0) {            // positive definite
    if (val(a) <= 0.0) { a2 = a; a3 = a; a4 = T(0.0); return; }
    if (fabs(val(a3) - val(a2)) < -val(a4)) {
      if (val(a) + 0.25 * (val(a3) - val(a2)) * (val(a3) - val(a2)) / val(a4) + val(a4) * (1.0 / 12.0) < 0.0) {
        if (val(a) < val(a3) && val(a) < val(a2)) { a3 = a; a2 = a; a4 = T(0.0); }
        else if (val(a3) > val(a2)) { a4 = 3.0 * (a2 - a); a3 = a2 - a4; }
        else { a4 = 3.0 * (a3 - a); a2 = a3 - a4; }
      }
    }
    return;
  }
  const bool flat = iv == 1 ? ((val(a) - val(a2)) * (val(a) - val(a3)) >= 0.0) : extm;
  if (flat) { a2 = a; a3 = a; a4 = T(0.0); return; }
  const double da1 = val(a3) - val(a2), da2 = da1 * da1, a6da = val(a4) * da1;
  if (a6da < -da2) { a4 = 3.0 * (a2 - a); a3 = a2 - a4; }
  else if (a6da > da2) { a4 = 3.0 * (a3 - a); a2 = a3 - a4; }
}
// q[0..K]: interface values of the cubic spline (modified in place by the large-scale constraints); out: A2, A3, A4 per layer.
// cs = true: cs_profile (no qmin tests; winds, w, delz), false: scalar_profile (T, tracers)
template <class T> DEV void limit_profile(const T* a, T* q, int K, int iv, int kord, double qmin, bool cs, T* A2, T* A3, T* A4) {
  T gm[KMAX]; bool extm[KMAX];
  const int ak = kord < 0 ? -kord : kord;
  q[1] = m_min(q[1], m_max(a[0], a[1])); q[1] = m_max(q[1], m_min(a[0], a[1]));
  gm[0] = T(0.0);
  for (int k = 1; k < K; k++) gm[k] = a[k] - a[k - 1];
  for (int k = 2; k <= K - 2; k++) {
    if (val(gm[k - 1]) * val(gm[k + 1]) > 0.0) { q[k] = m_min(q[k], m_max(a[k - 1], a[k])); q[k] = m_max(q[k], m_min(a[k - 1], a[k])); }
    else if (val(gm[k - 1]) > 0.0) q[k] = m_max(q[k], m_min(a[k - 1], a[k]));
    else { q[k] = m_min(q[k], m_max(a[k - 1], a[k])); if (iv == 0) q[k] = m_max(T(0.0), q[k]); }
  }
  q[K - 1] = m_min(q[K - 1], m_max(a[K - 2], a[K - 1])); q[K - 1] = m_max(q[K - 1], m_min(a[K - 2], a[K - 1]));
  for (int k = 0; k < K; k++) { A2[k] = q[k]; A3[k] = q[k + 1]; }
  for (int k = 0; k < K; k++)
    extm[k] = (k == 0 || k == K - 1) ? ((val(A2[k]) - val(a[k])) * (val(A3[k]) - val(a[k])) > 0.0) : (val(gm[k]) * val(gm[k + 1]) < 0.0);
  auto A4of = [&](int k) { return 3.0 * (2.0 * a[k] - (A2[k] + A3[k])); };
  auto A4of6 = [&](int k) { return 6.0 * a[k] - 3.0 * (A2[k] + A3[k]); };
  auto flat = [&](int k) { A2[k] = a[k]; A3[k] = a[k]; A4[k] = T(0.0); };
  auto huynh = [&](int k) {      // Huynh's second constraint on both edges
    T pmp_1 = a[k] - 2.0 * gm[k + 1], lac_1 = pmp_1 + 1.5 * gm[k + 2];
    A2[k] = m_min(m_max(A2[k], mn3(a[k], pmp_1, lac_1)), mx3(a[k], pmp_1, lac_1));
    T pmp_2 = a[k] + 2.0 * gm[k], lac_2 = pmp_2 - 1.5 * gm[k - 1];
    A3[k] = m_min(m_max(A3[k], mn3(a[k], pmp_2, lac_2)), mx3(a[k], pmp_2, lac_2));
  };
  // top two layers
  if (iv == 0) A2[0] = m_max(T(0.0), A2[0]);
  else if (iv == -1) { if (val(A2[0]) * val(a[0]) <= 0.0) A2[0] = T(0.0); }
  A4[0] = A4of(0); cs_limiters(extm[0], a[0], A2[0], A3[0], A4[0], 1);
  A4[1] = A4of(1); cs_limiters(extm[1], a[1], A2[1], A3[1], A4[1], 2);
  // interior
  for (int k = 2; k <= K - 3; k++) {
    const bool small = !cs && val(a[k]) < qmin;
    if (ak < 9) { huynh(k); A4[k] = A4of(k); }
    else if (ak == 9) {
      if (extm[k] && (extm[k - 1] || extm[k + 1] || small)) flat(k);
      else {
        A4[k] = cs ? A4of6(k) : A4of(k);
        if (fabs(val(A4[k])) > fabs(val(A2[k]) - val(A3[k]))) { huynh(k); A4[k] = cs ? A4of6(k) : A4of(k); }
      }
    } else if (ak == 10) {
      if (extm[k]) { if (small || extm[k - 1] || extm[k + 1]) flat(k); else A4[k] = A4of6(k); }
      else { A4[k] = A4of6(k); if (fabs(val(A4[k])) > fabs(val(A2[k]) - val(A3[k]))) { huynh(k); A4[k] = A4of6(k); } }
    } else if (ak == 12) {
      if (extm[k]) flat(k);
      else { A4[k] = A4of6(k); if (fabs(val(A4[k])) > fabs(val(A2[k]) - val(A3[k]))) { huynh(k); A4[k] = A4of6(k); } }
    } else if (ak == 13) {
      if (extm[k]) { if (extm[k - 1] && extm[k + 1]) flat(k); else { huynh(k); A4[k] = A4of(k); } }
      else A4[k] = A4of(k);
    } else if (ak == 14) A4[k] = A4of(k);
    else {                  // 11
      if (extm[k] && (extm[k - 1] || extm[k + 1] || small)) flat(k); else A4[k] = A4of(k);
    }
    if (iv == 0) cs_limiters(extm[k], a[k], A2[k], A3[k], A4[k], 0);
  }
  // bottom two layers
  if (iv == 0) A3[K - 1] = m_max(T(0.0), A3[K - 1]);
  else if (iv == -1) { if (val(A3[K - 1]) * val(a[K - 1]) <= 0.0) A3[K - 1] = T(0.0); }
  A4[K - 2] = A4of(K - 2); cs_limiters(extm[K - 2], a[K - 2], A2[K - 2], A3[K - 2], A4[K - 2], 2);
  A4[K - 1] = A4of(K - 1); cs_limiters(extm[K - 1], a[K - 1], A2[K - 1], A3[K - 1], A4[K - 1], 1);
}
}  // namespace rmp

// Remap one field with the nonlinear model's kord (8 .. 14; > 16 falls back to the unlimited profile).  Same inputs as S_remap.
struct S_remap_nl {
  static constexpr int NI = 5, NO = 1;
  struct P { int K, iv, use_dp2; int i0, i1, j0, j1; int kord, cs; double qmin; };
  template <class X> DEV static void eval(X& x, const P& p) {
    using T = typename X::T;
    if (!x.in_rect(p.i0, p.i1, p.j0, p.j1)) return;
    const int K = p.K;
    T a[KMAX], pe1[KMAX + 1], pe2[KMAX + 1], dp[KMAX], q[KMAX + 1], gam[KMAX + 1], A2[KMAX], A3[KMAX], A4[KMAX];
    double pe1v[KMAX + 1];
    for (int k = 0; k <= K; k++) { pe1[k] = x.in(1, k); pe2[k] = x.in(2, k); pe1v[k] = val(pe1[k]); }
    for (int k = 0; k < K; k++) { a[k] = x.in(0, k); dp[k] = pe1[k + 1] - pe1[k]; }
    T qs = (p.iv == -2) ? x.in(3, 0) : T(0.0);
    rmp::cs_profile(a, dp, q, gam, K, p.iv, qs);
    const int ak = p.kord < 0 ? -p.kord : p.kord;
    if (ak > 16) for (int k = 0; k < K; k++) { A2[k] = q[k]; A3[k] = q[k + 1]; A4[k] = 3.0 * (2.0 * a[k] - (A2[k] + A3[k])); }
    else rmp::limit_profile(a, q, K, p.iv, p.kord, p.qmin, p.cs != 0, A2, A3, A4);
    int k0 = 0;
    for (int k = 0; k < K; k++) {
      int l = rmp::find_layer(pe1v, val(pe2[k]), k0, K);
      T a2 = A2[l], a3 = A3[l], a4 = A4[l];
      T pl = (pe2[k] - pe1[l]) / dp[l];
      T res;
      if (val(pe2[k + 1]) <= pe1v[l + 1]) {
        T pr = (pe2[k + 1] - pe1[l]) / dp[l];
        res = a2 + 0.5 * (a4 + a3 - a2) * (pr + pl) - a4 * rmp::r3 * (pr * (pr + pl) + pl * pl);
        k0 = l;
      } else {
        T qsum = (pe1[l + 1] - pe2[k]) * (a2 + 0.5 * (a4 + a3 - a2) * (1.0 + pl) - a4 * (rmp::r3 * (1.0 + pl * (1.0 + pl))));
        for (int m = l + 1; m < K; m++) {
          if (val(pe2[k + 1]) > pe1v[m + 1]) qsum = qsum + dp[m] * a[m];
          else {
            T dpl = pe2[k + 1] - pe1[m];
            T esl = dpl / dp[m];
            qsum = qsum + dpl * (A2[m] + 0.5 * esl * (A3[m] - A2[m] + A4[m] * (1.0 - rmp::r23 * esl)));
            k0 = m;
            break;
          }
        }
        T den = p.use_dp2 ? x.in(4, k) : pe2[k + 1] - pe2[k];
        res = qsum / den;
      }
      x.out(0, k, res);
    }
  }
  template <class X> DEV static void eval_ad(X&, const P&) {}    // never active (detached inputs)
};


}  // namespace fv3lm
