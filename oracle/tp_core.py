"""ORACLE (test infrastructure only).  fv_tp_2d / xppm / yppm / copy_corners restated in
torch float64 from the reference's TL-module primal

  model_tlmadm/tp_core_tlm.F90   FV_TP_2D :83 (TLM :2123), XPPM :237 (TLM :2328),
                                 YPPM :957 (TLM :2496), COPY_CORNERS :2046
  (hand-written original: model/tp_core_nlm.F90:78-420)

The code is differentiable: the TL oracle is torch.func.jvp of these functions and the
AD oracle is torch.func.vjp -- exactly what Tapenade's tangent / reverse modes compute
for the same primal.  Only the linear orders the reference TL/AD
implement are restated (iord = 1, 2; tp_core_tlm.F90:2431-2466).

parity unpinned: the reference holds no test vectors for this routine.

Arrays: [6, K, NY, NX], Fortran (i, j) at [..., j+2, i+2].  Whole-tile decomposition:
is = js = 1, ie = je = N, npx = npy = N + 1.
"""
import torch
from .cubed_sphere import R, NG, copy_corners

P1, P2 = 7.0 / 12.0, -1.0 / 12.0
C1, C2, C3 = -2.0 / 14.0, 11.0 / 14.0, 5.0 / 14.0


class Grid:
    """metrics as torch tensors broadcastable against [6, K, NY, NX]"""

    def __init__(self, M):
        self.N = M["N"]; self.npx = M["npx"]; self.npy = M["npy"]; self.ng = NG
        self.is_ = 1; self.ie = self.N; self.js = 1; self.je = self.N
        self.isd = 1 - NG; self.ied = self.N + NG; self.jsd = 1 - NG; self.jed = self.N + NG
        for k, v in M.items():
            if hasattr(v, "ndim") and v.ndim >= 2:
                t = torch.as_tensor(v, dtype=torch.float64)
                if v.ndim == 3:
                    t = t[:, None]            # [6, 1, NY, NX]
                elif v.ndim == 4:
                    t = t[:, None]            # [6, 1, NY, NX, c]
                elif v.ndim == 2:
                    t = t[:, None]            # [6, 1, NX]   (edge factors)
                setattr(self, k, t)
            else:
                setattr(self, k, v)


def sh(a, js, is_):
    """a[..., js, is_] with Fortran inclusive ranges given as (lo, hi) tuples"""
    return a[..., R(js[0], js[1]), R(is_[0], is_[1])]


def _al_x(q, dxa, npx, i0, i1, j0, j1):
    """al(i) for i in [0, npx+1] (whole tile), rows j0..j1.  Returns tensor indexed so
    that al[..., i - 0] corresponds to Fortran i = 0 .. npx+1."""
    def Q(a, b):  # q(a..b, j0..j1)
        return q[..., R(j0, j1), R(a, b)]
    def D(a, b):
        return dxa[..., R(j0, j1), R(a, b)]
    lo, hi = 0, npx + 1
    al = P1 * (Q(lo - 1, hi - 1) + Q(lo, hi)) + P2 * (Q(lo - 2, hi - 2) + Q(lo + 1, hi + 1))
    cols = list(torch.unbind(al, dim=-1))
    def q1(i): return q[..., R(j0, j1), i + NG - 1]
    def d1(i): return dxa[..., R(j0, j1), i + NG - 1]
    def edge(i):
        return 0.5 * (((2. * d1(i - 1) + d1(i - 2)) * q1(i - 1) - d1(i - 1) * q1(i - 2)) / (d1(i - 2) + d1(i - 1)) +
                      ((2. * d1(i) + d1(i + 1)) * q1(i) - d1(i) * q1(i + 1)) / (d1(i) + d1(i + 1)))
    cols[0] = C1 * q1(-2) + C2 * q1(-1) + C3 * q1(0)
    cols[1] = edge(1)
    cols[2] = C3 * q1(1) + C2 * q1(2) + C1 * q1(3)
    cols[npx - 1] = C1 * q1(npx - 3) + C2 * q1(npx - 2) + C3 * q1(npx - 1)
    cols[npx] = edge(npx)
    cols[npx + 1] = C3 * q1(npx) + C2 * q1(npx + 1) + C1 * q1(npx + 2)
    return torch.stack(cols, dim=-1)     # [..., nj, npx+2]


def xppm(q, c, iord, g, j0, j1):
    """flux(is:ie+1, j0:j1).  q, c: full arrays.  iord: python int or per-level list (len K).
    Returns a full-size array with the flux stored on its range (zero elsewhere)."""
    npx = g.npx
    is_, ie = g.is_, g.ie
    out = torch.zeros_like(q)
    cc = c[..., R(j0, j1), R(is_, ie + 1)]
    qm = q[..., R(j0, j1), R(is_ - 1, ie)]
    qp = q[..., R(j0, j1), R(is_, ie + 1)]
    up = torch.where(cc > 0., qm, qp)
    al = _al_x(q, g.dxa, npx, is_, ie, j0, j1)           # index = Fortran i (0..npx+1)
    al_m = al[..., is_ - 1: ie + 1]                      # al(i-1)
    al_0 = al[..., is_: ie + 2]
    al_p = al[..., is_ + 1: ie + 3]
    f_pos = qm + (1. - cc) * (al_0 - qm - cc * (al_m + al_0 - (qm + qm)))
    f_neg = qp + (1. + cc) * (al_0 - qp + cc * (al_0 + al_p - (qp + qp)))
    f2 = torch.where(cc > 0., f_pos, f_neg)
    if isinstance(iord, int):
        fl = up if iord == 1 else f2
    else:
        sel = torch.tensor([o == 1 for o in iord], dtype=torch.bool).view(1, -1, 1, 1)
        fl = torch.where(sel, up, f2)
    out[..., R(j0, j1), R(is_, ie + 1)] = fl
    return out


def _T(a):
    return a.transpose(-1, -2)


def yppm(q, c, jord, g, i0, i1):
    """flux(i0:i1, js:je+1): the y sweep is the x sweep on transposed slabs (the reference
    code is the literal transpose with dya in place of dxa, tp_core_tlm.F90:2496-2670)."""
    class GT:
        pass
    gt = GT()
    gt.npx = g.npy; gt.is_ = g.js; gt.ie = g.je; gt.dxa = _T(g.dya)
    return _T(xppm(_T(q), _T(c), jord, gt, i0, i1))


def fv_tp_2d(q, crx, cry, hord, xfx, yfx, g, ra_x, ra_y, mfx=None, mfy=None):
    """returns fx(is:ie+1, js:je), fy(is:ie, js:je+1) as full arrays, and q with its corner
    ghost cells as the routine leaves them (copy_corners dir 1)."""
    is_, ie, js, je = g.is_, g.ie, g.js, g.je
    isd, ied, jsd, jed = g.isd, g.ied, g.jsd, g.jed
    npx, npy = g.npx, g.npy
    A = g.area
    q = copy_corners(q, npx, npy, 2)
    fy2 = yppm(q, cry, hord, g, isd, ied)
    fyy = yfx * fy2
    q_i = torch.zeros_like(q)
    r = (..., R(js, je), R(isd, ied)); rp = (..., R(js + 1, je + 1), R(isd, ied))
    q_i[r] = (q[r] * A[r] + fyy[r] - fyy[rp]) / ra_y[r]
    fx = xppm(q_i, crx, hord, g, js, je)
    q = copy_corners(q, npx, npy, 1)
    fx2 = xppm(q, crx, hord, g, jsd, jed)
    fx1 = xfx * fx2
    q_j = torch.zeros_like(q)
    r = (..., R(jsd, jed), R(is_, ie)); rp = (..., R(jsd, jed), R(is_ + 1, ie + 1))
    q_j[r] = (q[r] * A[r] + fx1[r] - fx1[rp]) / ra_x[r]
    fy = yppm(q_j, cry, hord, g, is_, ie)
    mx = xfx if mfx is None else mfx
    my = yfx if mfy is None else mfy
    fxo = torch.zeros_like(q); fyo = torch.zeros_like(q)
    r = (..., R(js, je), R(is_, ie + 1))
    fxo[r] = 0.5 * (fx[r] + fx2[r]) * mx[r]
    r = (..., R(js, je + 1), R(is_, ie))
    fyo[r] = 0.5 * (fy[r] + fy2[r]) * my[r]
    return fxo, fyo, q
