// Linearised boundary-layer turbulence (see turb.h).  Three kernels, all one thread per column with i fastest across the warp, so
// every level access of a warp is one contiguous 256-byte segment.  HBM bound: the solve reads 9 coefficient arrays + pk and
// reads / writes 7 fields = 24 array passes of algorithmic traffic per call; the two-sweep form re-reads each field once more
// (the down-sweep result is parked in the field itself instead of a per-thread work array, which would spill to local memory).
// Fields that share a coefficient set are solved by the same thread so that the set is loaded once for all of them.
#include "engine.h"
#include "turb.h"
#include <math.h>

namespace fv3lm {

namespace {

// vtrilupert (:562-579)
struct KTurbLU {
  TurbDims s; double* a[3]; double* b[3]; const double* c[3];
  DEV void operator()(int i, int j, int z) const {
    const int set = z % 3, sub = z / 3;
    const size_t ls = (size_t)s.nx * s.ny, o = (size_t)sub * s.K * ls + (size_t)j * s.nx + i;
    double* A = a[set] + o; double* B = b[set] + o; const double* C = c[set] + o;
    double bp = 1.0 / B[0];
    B[0] = bp;
    for (int l = 1; l < s.K; l++) {
      const double al = A[l * ls] * bp;
      A[l * ls] = al;
      bp = 1.0 / (B[l * ls] - C[(l - 1) * ls] * al);
      B[l * ls] = bp;
    }
  }
};

// compute_pressures (utils:359-391), pk only
struct KTurbPk {
  TurbDims s; const double* delp; double* pk; double ptop, kappa;
  DEV void operator()(int i, int j, int z) const {
    const size_t ls = (size_t)s.nx * s.ny, o = (size_t)z * s.K * ls + (size_t)j * s.nx + i;
    double pe = ptop, pek = pow(pe, kappa), lpe = log(pe);
    for (int l = 0; l < s.K; l++) {
      const double pe1 = pe + LDG(delp + o + l * ls), pek1 = pow(pe1, kappa), lpe1 = log(pe1);
      pk[o + l * ls] = (pek1 - pek) / (kappa * (lpe1 - lpe));
      pe = pe1; pek = pek1; lpe = lpe1;
    }
  }
};

// vtrisolvepert (:583-675) for the fields of one coefficient set.  group 0: u v (akv) ; 1: t (aks, with the t <-> theta scaling
// of step_tl :257 / :269 and its adjoint :327 / :339) ; 2: qv qi ql o3 (akq; qv with ygswitch = 1, the tracers with 0)
struct KTurbSolve {
  TurbDims s; const double* a[3]; const double* b[3]; const double* c[3]; const double* pk;
  double* y[3][4]; double p0k; int adjoint;
  DEV void operator()(int i, int j, int z) const {
    const int grp = z % 3, sub = z / 3, K = s.K;
    const int nf = grp == 0 ? 2 : grp == 1 ? 1 : 4;
    const bool sc = grp == 1;
    const size_t ls = (size_t)s.nx * s.ny, o = (size_t)sub * K * ls + (size_t)j * s.nx + i;
    const double* A = a[grp] + o; const double* B = b[grp] + o; const double* C = c[grp] + o; const double* PK = pk + o;
    double* Y[4]; for (int m = 0; m < nf; m++) Y[m] = y[grp][m] + o;
    auto yg = [&](int m) { return grp != 2 || m == 0; };            // ygswitch = 1: winds, temperature, qv
    double cur[4];
    if (!adjoint) {
      // sweep down, modifying the right-hand side with the multiplier a (:604-606)
      for (int m = 0; m < nf; m++) {
        double v = Y[m][0];
        if (sc) { v = p0k * v / LDG(PK); Y[m][0] = v; }
        cur[m] = v;
      }
      for (int l = 1; l < K; l++) {
        const double al = LDG(A + l * ls);
        for (int m = 0; m < nf; m++) {
          double v = Y[m][l * ls];
          if (sc) v = p0k * v / LDG(PK + l * ls);
          v = v - al * cur[m];
          if (l < K - 1) Y[m][l * ls] = v;                            // (level K-1 stays in the register)
          cur[m] = v;
        }
      }
      // surface level (:609-613), then sweep up; b holds the inverse of the main diagonal (:615-617)
      {
        const double bl = LDG(B + (K - 1) * ls), bm = LDG(B + (K - 2) * ls), am = LDG(A + (K - 1) * ls), cm = LDG(C + (K - 2) * ls);
        for (int m = 0; m < nf; m++) {
          const double v = yg(m) ? cur[m] * bl : cur[m] * bm / (bm - am * (1.0 + cm * bm));
          cur[m] = v;
          Y[m][(K - 1) * ls] = sc ? LDG(PK + (K - 1) * ls) * v / p0k : v;
        }
      }
      for (int l = K - 2; l >= 0; l--) {
        const double bl = LDG(B + l * ls), cl = LDG(C + l * ls);
        for (int m = 0; m < nf; m++) {
          const double v = bl * (Y[m][l * ls] - cl * cur[m]);
          cur[m] = v;
          Y[m][l * ls] = sc ? LDG(PK + l * ls) * v / p0k : v;
        }
      }
    } else {
      // ygswitch = 1: (LU)' = U'L' (:623-633).  ygswitch = 0: line-by-line adjoint of the sweep up, the surface fix and the
      // sweep down (:637-651).  Both march down with b, c and back up with a, so they share the loads.
      for (int m = 0; m < nf; m++) {
        double v = Y[m][0];
        if (sc) v = LDG(PK) * v / p0k;
        cur[m] = v;
      }
      {
        const double b0 = LDG(B);
        for (int m = 0; m < nf; m++) if (yg(m)) cur[m] = cur[m] * b0;
      }
      for (int l = 1; l < K; l++) {
        const double bl = LDG(B + l * ls), bp = LDG(B + (l - 1) * ls), cp = LDG(C + (l - 1) * ls);
        for (int m = 0; m < nf; m++) {
          double v = Y[m][l * ls];
          if (sc) v = LDG(PK + l * ls) * v / p0k;
          if (yg(m)) {
            Y[m][(l - 1) * ls] = cur[m];
            cur[m] = bl * (v - cp * cur[m]);
          } else {
            v = v - cp * bp * cur[m];
            Y[m][(l - 1) * ls] = bp * cur[m];
            cur[m] = v;
          }
        }
      }
      {
        const double bm = LDG(B + (K - 2) * ls), am = LDG(A + (K - 1) * ls), cm = LDG(C + (K - 2) * ls);
        for (int m = 0; m < nf; m++) {
          if (!yg(m)) cur[m] = bm * cur[m] / (bm - am * (cm * bm + 1.0));
          Y[m][(K - 1) * ls] = sc ? p0k * cur[m] / LDG(PK + (K - 1) * ls) : cur[m];
        }
      }
      for (int l = K - 2; l >= 0; l--) {
        const double an = LDG(A + (l + 1) * ls);
        for (int m = 0; m < nf; m++) {
          const double v = Y[m][l * ls] - an * cur[m];
          cur[m] = v;
          Y[m][l * ls] = sc ? p0k * v / LDG(PK + l * ls) : v;
        }
      }
    }
  }
};

}  // namespace

void turb_lu(const TurbDims& s, TurbLtraj& lt) {
  KTurbLU k{s, {lt.d[TURB_AKV], lt.d[TURB_AKS], lt.d[TURB_AKQ]}, {lt.d[TURB_BKV], lt.d[TURB_BKS], lt.d[TURB_BKQ]},
            {lt.d[TURB_CKV], lt.d[TURB_CKS], lt.d[TURB_CKQ]}};
  launch3d(k, s.nx, s.ny, s.nsub * 3);
}

void turb_pk(const TurbDims& s, const double* delp, double* pk, double ptop, double kappa) {
  launch3d(KTurbPk{s, delp, pk, ptop, kappa}, s.nx, s.ny, s.nsub);
}

void turb_solve(const TurbDims& s, const TurbLtraj& lt, double* const f[7], double p0k, bool adjoint) {
  KTurbSolve k{s, {lt.d[TURB_AKV], lt.d[TURB_AKS], lt.d[TURB_AKQ]}, {lt.d[TURB_BKV], lt.d[TURB_BKS], lt.d[TURB_BKQ]},
               {lt.d[TURB_CKV], lt.d[TURB_CKS], lt.d[TURB_CKQ]}, lt.d[TURB_PK],
               {{f[0], f[1], nullptr, nullptr}, {f[2], nullptr, nullptr, nullptr}, {f[3], f[4], f[5], f[6]}}, p0k, adjoint ? 1 : 0};
  launch3d(k, s.nx, s.ny, s.nsub * 3);
}

}  // namespace fv3lm
