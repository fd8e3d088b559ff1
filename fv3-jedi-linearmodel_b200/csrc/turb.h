// Linearised boundary-layer turbulence of fv3jedi_lm:
// src/physics/turbulence/fv3jedi_lm_turbulence_mod.F90  step_nl :149-213, step_tl :218-281, step_ad :285-348,
// vtrilupert :562-579, vtrisolvepert :583-675; pk from src/utils/fv3jedi_lm_utils_mod.F90:359-391.
// All arrays are compact [nsub][K][ny][nx] (the API layout); columns are independent, nothing is exchanged.
#pragma once
#include <stddef.h>

namespace fv3lm {

enum { TURB_AKV = 0, TURB_BKV, TURB_CKV, TURB_AKS, TURB_BKS, TURB_CKS, TURB_AKQ, TURB_BKQ, TURB_CKQ, TURB_PK, TURB_NARR };

// local trajectory of one time level (local_traj_turbulence, :30-36): LU-decomposed diagonals + p^kappa, device resident
struct TurbLtraj {
  double* d[TURB_NARR] = {};
  bool set = false;
};

struct TurbDims { int nx, ny, nsub, K; };

// b <- 1 / pivot, a <- multiplier (vtrilupert); one thread per column and coefficient set
void turb_lu(const TurbDims& s, TurbLtraj& lt);
// pk(k) = (pe(k)^kappa - pe(k-1)^kappa) / (kappa (ln pe(k) - ln pe(k-1))),  pe(0) = ptop, pe(k) = pe(k-1) + delp(k)
void turb_pk(const TurbDims& s, const double* delp, double* pk, double ptop, double kappa);
// the seven solves of step_tl / step_nl (adjoint = false) or step_ad (adjoint = true), in place.
// fields: u v t qv qi ql o3 (compact device arrays)
void turb_solve(const TurbDims& s, const TurbLtraj& lt, double* const fields[7], double p0k, bool adjoint);

}  // namespace fv3lm
