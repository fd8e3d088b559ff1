"""ORACLE (test infrastructure only).  Linearised boundary-layer turbulence of fv3jedi_lm, restated in numpy float64 from

  src/physics/turbulence/fv3jedi_lm_turbulence_mod.F90
      step_nl :149-213, step_tl :218-281, step_ad :285-348   (t -> theta, seven tridiagonal solves, theta -> t)
      vtrilupert :562-579      LU decomposition of the tridiagonal systems of the trajectory
      vtrisolvepert :583-675   solve with the decomposed system: phase 1 (TLM), phase 2 (ADM), ygswitch 1 / 0
  src/utils/fv3jedi_lm_utils_mod.F90:359-391  compute_pressures (pk)
  src/utils/fv3jedi_lm_const_mod.F90:43       p00 = 1e5

Arrays are [..., K, ny, nx] with the level axis third from last (the compact API layout); loops run over levels only, like the
reference's array-section statements.  BL_DRIVER (bldriver.F90), which produces the diagonals from the trajectory, is outside the
path (trajectory set-up on the caller's side of the boundary).  parity unpinned (the reference ships no vectors); what pins this
file: the decomposed solve reproduces a dense solve of the original system (ygswitch = 1), and phase 2 is the exact transpose of
phase 1 for both switches (tests/test_turbulence.py).
"""
import numpy as np

P00 = 100000.0


def L(a, l):
    """level l (0-based) of an array [..., K, ny, nx]"""
    return a[..., l, :, :]


def compute_pk(delp, ptop, kappa):
    """utils:359-391 -- p^kappa at the mid-points from the edge pressures"""
    K = delp.shape[-3]
    pe = np.empty(delp.shape[:-3] + (K + 1,) + delp.shape[-2:])
    pe[..., 0, :, :] = ptop
    for l in range(1, K + 1):
        pe[..., l, :, :] = pe[..., l - 1, :, :] + delp[..., l - 1, :, :]
    lpe = np.log(pe)
    pek = pe ** kappa
    return (pek[..., 1:, :, :] - pek[..., :-1, :, :]) / (kappa * (lpe[..., 1:, :, :] - lpe[..., :-1, :, :]))


def vtrilupert(a, b, c):
    """:562-579 -- returns the decomposed (a, b): a = multipliers, b = inverse pivots"""
    a = a.copy(); b = b.copy()
    K = a.shape[-3]
    b[..., 0, :, :] = 1.0 / L(b, 0)
    for l in range(1, K):
        a[..., l, :, :] = L(a, l) * L(b, l - 1)
        b[..., l, :, :] = 1.0 / (L(b, l) - L(c, l - 1) * L(a, l))
    return a, b


def vtrisolvepert(a, b, c, y, phase, ygswitch):
    """:583-675 -- a, b decomposed.  Returns the new y."""
    y = y.copy()
    K = y.shape[-3]
    lm = K - 1
    if phase == 1:
        for l in range(1, K):                                          # :604-606
            y[..., l, :, :] = L(y, l) - L(a, l) * L(y, l - 1)
        if ygswitch == 1:                                              # :609-613
            y[..., lm, :, :] = L(y, lm) * L(b, lm)
        else:
            y[..., lm, :, :] = L(y, lm) * L(b, lm - 1) / (L(b, lm - 1) - L(a, lm) * (1.0 + L(c, lm - 1) * L(b, lm - 1)))
        for l in range(lm - 1, -1, -1):                                # :615-617
            y[..., l, :, :] = L(b, l) * (L(y, l) - L(c, l) * L(y, l + 1))
    elif ygswitch == 1:                                                # :623-633
        y[..., 0, :, :] = L(y, 0) * L(b, 0)
        for l in range(1, K):
            y[..., l, :, :] = L(b, l) * (L(y, l) - L(c, l - 1) * L(y, l - 1))
        for l in range(lm - 1, -1, -1):
            y[..., l, :, :] = L(y, l) - L(a, l + 1) * L(y, l + 1)
    else:                                                              # :637-651
        for l in range(0, lm):
            y[..., l + 1, :, :] = L(y, l + 1) - L(c, l) * L(b, l) * L(y, l)
            y[..., l, :, :] = L(b, l) * L(y, l)
        y[..., lm, :, :] = L(b, lm - 1) * L(y, lm) / (L(b, lm - 1) - L(a, lm) * (L(c, lm - 1) * L(b, lm - 1) + 1.0))
        for l in range(lm, 0, -1):
            y[..., l - 1, :, :] = L(y, l - 1) - L(a, l) * L(y, l)
    return y


def set_ltraj(co, delp, ptop, kappa):
    """the part of set_ltraj :375-533 after BL_DRIVER: pk and the three decompositions.  co: dict akv ... ckq as BL_DRIVER returns"""
    lt = dict(pk=compute_pk(delp, ptop, kappa))
    for s in "vsq":
        a, b = vtrilupert(co["ak" + s], co["bk" + s], co["ck" + s])
        lt["ak" + s], lt["bk" + s], lt["ck" + s] = a, b, co["ck" + s]
    return lt


def step(lt, x, kappa, phase):
    """step_tl (phase 1; step_nl applies the same to the trajectory) / step_ad (phase 2).  x: dict u v t qv qi ql o3 (+ others,
    passed through).  Returns a new dict."""
    o = {k: v.copy() for k, v in x.items()}
    p0k = P00 ** kappa
    if phase == 1:
        o["t"] = p0k * o["t"] / lt["pk"]                                # :257
    else:
        o["t"] = lt["pk"] * o["t"] / p0k                                # :327
    o["u"] = vtrisolvepert(lt["akv"], lt["bkv"], lt["ckv"], o["u"], phase, 1)
    o["v"] = vtrisolvepert(lt["akv"], lt["bkv"], lt["ckv"], o["v"], phase, 1)
    o["t"] = vtrisolvepert(lt["aks"], lt["bks"], lt["cks"], o["t"], phase, 1)
    o["qv"] = vtrisolvepert(lt["akq"], lt["bkq"], lt["ckq"], o["qv"], phase, 1)
    for n in ("qi", "ql", "o3"):
        o[n] = vtrisolvepert(lt["akq"], lt["bkq"], lt["ckq"], o[n], phase, 0)
    if phase == 1:
        o["t"] = lt["pk"] * o["t"] / p0k                                # :269
    else:
        o["t"] = p0k * o["t"] / lt["pk"]                                # :339
    return o
