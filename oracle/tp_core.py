"""ORACLE (test infrastructure only).  fv_tp_2d / xppm / yppm / copy_corners restated in
torch float64 from the reference's TL-module primal

  model_tlmadm/tp_core_tlm.F90   FV_TP_2D :83 (TLM :2123), XPPM :237 (TLM :2328),
                                 YPPM :957 (TLM :2496), COPY_CORNERS :2046
  (hand-written original: model/tp_core_nlm.F90:78-420)

The code is differentiable: the TL oracle is torch.func.jvp of these functions and the
AD oracle is torch.func.vjp -- exactly what Tapenade's tangent / reverse modes compute
for the same primal.  Only the linear orders the reference TL/AD
implement are restated (iord = 1, 2, 333; tp_core_tlm.F90:2431-2488).

parity: pinned to the reference's own tangent code (tests/test_ref_tlm.py: XPPM / YPPM / FV_TP_2D / DELN_FLUX / COPY_CORNERS _TLM
transliterated; tests/test_ref_golden.py: the reference's D_SW_TLM / D_SW_BWD / DYN_CORE_TLM executed, which call these routines).

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


S11, S14, S15, R3 = 11. / 14., 4. / 7., 3. / 14., 1. / 3.


def _sign(a, b):
    """Fortran SIGN(a, b)"""
    return torch.where(b >= 0., a.abs(), -a.abs())


def _max3(a, b, c):
    return torch.maximum(torch.maximum(a, b), c)


def _min3(a, b, c):
    return torch.minimum(torch.minimum(a, b), c)


def pert_ppm(a0, al, ar, iv):
    """tp_core_nlm.F90:953-1012"""
    z = torch.zeros_like(al)
    if iv == 0:
        a4 = -3. * (ar + al); da1 = ar - al
        fmin = a0 + 0.25 / torch.where(a4 == 0., torch.ones_like(a4), a4) * da1 ** 2 + a4 * (1. / 12.)
        act = (da1.abs() < -a4) & (fmin < 0.)
        both = (ar > 0.) & (al > 0.)
        ar_n = torch.where(act, torch.where(both, z, torch.where(da1 > 0., -2. * al, ar)), ar)
        al_n = torch.where(act, torch.where(both, z, torch.where(da1 > 0., al, -2. * ar)), al)
        neg = a0 <= 0.
        return torch.where(neg, z, al_n), torch.where(neg, z, ar_n)
    da1 = al - ar; da2 = da1 ** 2; a6da = 3. * (al + ar) * da1
    ar_n = torch.where(a6da < -da2, -2. * al, ar)
    al_n = torch.where((a6da >= -da2) & (a6da > da2), -2. * ar, al)
    opp = al * ar < 0.
    return torch.where(opp, al_n, z), torch.where(opp, ar_n, z)


def _mono_blbr_x(q, dxa, npx, j0, j1, iord):
    """monotone PPM (iord 8..13) in perturbation form, tp_core_nlm.F90:470-569: bl, br of the cells i = 0 .. npx (last dim index = i)"""
    near_zero, ppm_fac = 1.e-25, 1.5
    def Q(a, b):
        return q[..., R(j0, j1), R(a, b)]
    def q1(i): return q[..., R(j0, j1), i + NG - 1]
    def d1(i): return dxa[..., R(j0, j1), i + NG - 1]
    def DM(a, b):       # dm(a..b)
        xt = 0.25 * (Q(a + 1, b + 1) - Q(a - 1, b - 1))
        c0 = Q(a, b)
        return _sign(_min3(xt.abs(), _max3(Q(a - 1, b - 1), c0, Q(a + 1, b + 1)) - c0, c0 - _min3(Q(a - 1, b - 1), c0, Q(a + 1, b + 1))), xt)
    lo, hi = 0, npx
    al = 0.5 * (Q(lo - 1, hi) + Q(lo, hi + 1)) + R3 * (DM(lo - 1, hi) - DM(lo, hi + 1))      # al(lo .. hi+1)
    q0 = Q(lo, hi)
    al0, al1 = al[..., :-1], al[..., 1:]
    dm0 = DM(lo, hi)
    if iord in (8, 11):
        xt = (2. if iord == 8 else ppm_fac) * dm0
        bl = -_sign(torch.minimum(xt.abs(), (al0 - q0).abs()), xt)
        br = _sign(torch.minimum(xt.abs(), (al1 - q0).abs()), xt)
    else:
        bl = al0 - q0; br = al1 - q0
        dq = lambda a, b: 2. * (Q(a + 1, b + 1) - Q(a, b))
        pmp_2 = dq(lo - 1, hi - 1); lac_2 = pmp_2 - 0.75 * dq(lo - 2, hi - 2)
        z = torch.zeros_like(bl)
        br_l = torch.minimum(_max3(z, pmp_2, lac_2), torch.maximum(br, _min3(z, pmp_2, lac_2)))
        pmp_1 = -dq(lo, hi); lac_1 = pmp_1 + 0.75 * dq(lo + 1, hi + 1)
        bl_l = torch.minimum(_max3(z, pmp_1, lac_1), torch.maximum(bl, _min3(z, pmp_1, lac_1)))
        flat = (DM(lo - 1, hi - 1).abs() + dm0.abs() + DM(lo + 1, hi + 1).abs()) < near_zero
        lim = (3. * (bl + br)).abs() > (bl - br).abs()
        bl, br = torch.where(flat, z, torch.where(lim, bl_l, bl)), torch.where(flat, z, torch.where(lim, br_l, br))
    if iord in (9, 13):
        bl, br = pert_ppm(q0, bl, br, 0)
    bl = list(torch.unbind(bl, -1)); br = list(torch.unbind(br, -1))
    dmv = list(torch.unbind(DM(-1, npx + 1), -1))
    dm = lambda i: dmv[i + 1]
    alv = lambda i: al[..., i - lo]
    def two_sided(i):   # edge between cells i-1 and i, limited by the four cells around it
        xt = 0.5 * (((2. * d1(i - 1) + d1(i - 2)) * q1(i - 1) - d1(i - 1) * q1(i - 2)) / (d1(i - 2) + d1(i - 1)) +
                    ((2. * d1(i) + d1(i + 1)) * q1(i) - d1(i) * q1(i + 1)) / (d1(i) + d1(i + 1)))
        xt = torch.maximum(xt, torch.minimum(torch.minimum(q1(i - 2), q1(i - 1)), torch.minimum(q1(i), q1(i + 1))))
        return torch.minimum(xt, torch.maximum(torch.maximum(q1(i - 2), q1(i - 1)), torch.maximum(q1(i), q1(i + 1))))
    # west edge
    bl[0] = S14 * dm(-1) + S11 * (q1(-1) - q1(0))
    xt = two_sided(1)
    br[0] = xt - q1(0); bl[1] = xt - q1(1)
    xt = S15 * q1(1) + S11 * q1(2) - S14 * dm(2)
    br[1] = xt - q1(1); bl[2] = xt - q1(2)
    br[2] = alv(3) - q1(2)
    # east edge
    bl[npx - 2] = alv(npx - 2) - q1(npx - 2)
    xt = S15 * q1(npx - 1) + S11 * q1(npx - 2) + S14 * dm(npx - 2)
    br[npx - 2] = xt - q1(npx - 2); bl[npx - 1] = xt - q1(npx - 1)
    xt = two_sided(npx)
    br[npx - 1] = xt - q1(npx - 1); bl[npx] = xt - q1(npx)
    br[npx] = S11 * (q1(npx + 1) - q1(npx)) - S14 * dm(npx + 1)
    for i in (0, 1, 2, npx - 2, npx - 1, npx):
        bl[i], br[i] = pert_ppm(q1(i), bl[i], br[i], 1)
    return torch.stack(bl, -1), torch.stack(br, -1)


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
    # iord = 333: perfectly linear third-order scheme, no cube-edge special cases (tp_core_tlm.F90:2467-2488)
    qmm = q[..., R(j0, j1), R(is_ - 2, ie - 1)]
    qpp = q[..., R(j0, j1), R(is_ + 1, ie + 2)]
    f3 = torch.where(cc > 0., (2.0 * qp + 5.0 * qm - qmm) / 6.0 - 0.5 * cc * (qp - qm) + cc * cc / 6.0 * (qp - 2.0 * qm + qmm),
                     (2.0 * qm + 5.0 * qp - qpp) / 6.0 - 0.5 * cc * (qp - qm) + cc * cc / 6.0 * (qpp - 2.0 * qp + qm))
    table = {1: up, 2: f2, 333: f3}
    for o in set([iord] if isinstance(iord, int) else iord):
        if 8 <= o <= 13:       # monotone schemes of the nonlinear model (trajectory side of a two-sided configuration)
            bl, br = _mono_blbr_x(q, g.dxa, npx, j0, j1, o)          # index = Fortran i (0..npx)
            blm, brm = bl[..., is_ - 1: ie + 1], br[..., is_ - 1: ie + 1]
            bl0, br0 = bl[..., is_: ie + 2], br[..., is_: ie + 2]
            table[o] = torch.where(cc > 0., qm + (1. - cc) * (brm - cc * (blm + brm)), qp + (1. + cc) * (bl0 + cc * (bl0 + br0)))
        if 3 <= o <= 7:        # smoothness-switch schemes on the unlimited edge values (tp_core_nlm.F90:327-467)
            alo = al
            if o == 7:         # positivity of the edge values: interior -> mean of the two cells, cube-edge values -> 0
                mean = 0.5 * (q[..., R(j0, j1), R(-1, npx)] + q[..., R(j0, j1), R(0, npx + 1)])
                ii = torch.arange(0, npx + 2)
                edge = ((ii <= 2) | (ii >= npx - 1)).view(1, 1, 1, -1)
                alo = torch.where(al < 0., torch.where(edge, torch.zeros_like(al), mean), al)
            qc = q[..., R(j0, j1), R(0, npx)]
            bl = alo[..., 0: npx + 1] - qc; br = alo[..., 1: npx + 2] - qc; b0 = bl + br
            sl = lambda a, d: a[..., is_ + d: ie + 2 + d]          # cell i + d for faces i = is .. ie+1
            blm, brm, b0m, blp, brp, b0p = sl(bl, -1), sl(br, -1), sl(b0, -1), sl(bl, 0), sl(br, 0), sl(b0, 0)
            if o in (3, 4):
                s5 = b0.abs() < (bl - br).abs(); s6 = 3. * b0.abs() < (bl - br).abs()
                s5m, s5p, s6m, s6p = sl(s5, -1), sl(s5, 0), sl(s6, -1), sl(s6, 0)
                z = torch.zeros_like(cc)
                if o == 3:
                    f1p = torch.where(s6m | s5p, brm - cc * b0m, torch.where(s5m, _sign(torch.minimum(blm.abs(), brm.abs()), brm), z))
                    f1n = torch.where(s6p | s5m, blp + cc * b0p, torch.where(s5p, _sign(torch.minimum(blp.abs(), brp.abs()), blp), z))
                    table[o] = torch.where(cc > 0., qm + (1. - cc.abs()) * f1p, qp + (1. - cc.abs()) * f1n)
                else:
                    table[o] = torch.where(cc > 0., qm + torch.where(s6m | s5p, (1. - cc) * (brm - cc * b0m), z),
                                           qp + torch.where(s6p | s5m, (1. + cc) * (blp + cc * b0p), z))
            else:
                s5 = (bl * br < 0.) if o == 5 else ((3. * b0).abs() < (bl - br).abs())
                on = sl(s5, -1) | sl(s5, 0)
                f1 = torch.where(cc > 0., (1. - cc) * (brm - cc * b0m), (1. + cc) * (blp + cc * b0p))
                table[o] = up + torch.where(on, f1, torch.zeros_like(f1))
    out[..., R(j0, j1), R(is_, ie + 1)] = select_ord(iord, table)
    return out


def select_ord(iord, table):
    """iord: python int or per-level list (len K); table: order -> tensor [6, K, ...]"""
    if isinstance(iord, int):
        return table[iord]
    fl = None
    for o in sorted(set(iord)):
        sel = torch.tensor([x == o for x in iord], dtype=torch.bool).view(1, -1, 1, 1)
        fl = table[o] if fl is None else torch.where(sel, table[o], fl)
    return fl


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


def deln_flux(nord, damp, q, fx, fy, g, mass=None):
    """model/tp_core_nlm.F90:1015-1163 (del-n damping flux added to fx, fy).  nord: python
    int (0..2); damp: float.  q is the transported field AFTER fv_tp_2d (corner state
    irrelevant: copy_corners is re-applied on d2)."""
    is_, ie, js, je = g.is_, g.ie, g.js, g.je
    npx, npy = g.npx, g.npy
    def S(a, i0, i1, j0, j1): return a[..., R(j0, j1), R(i0, i1)]
    def put(a, i0, i1, j0, j1, v):
        a = a.clone(); a[..., R(j0, j1), R(i0, i1)] = v; return a
    i1_, i2_, j1_, j2_ = is_ - 1 - nord, ie + 1 + nord, js - 1 - nord, je + 1 + nord
    d2 = torch.zeros_like(q)
    d2 = put(d2, i1_, i2_, j1_, j2_, (S(q, i1_, i2_, j1_, j2_) if mass is not None else damp * S(q, i1_, i2_, j1_, j2_)))
    fx2 = torch.zeros_like(q); fy2 = torch.zeros_like(q)
    if nord > 0:
        d2 = copy_corners(d2, npx, npy, 1)
    i0, i1, j0, j1 = is_ - nord, ie + nord + 1, js - nord, je + nord
    fx2 = put(fx2, i0, i1, j0, j1, S(g.del6_v, i0, i1, j0, j1) * (S(d2, i0 - 1, i1 - 1, j0, j1) - S(d2, i0, i1, j0, j1)))
    if nord > 0:
        d2 = copy_corners(d2, npx, npy, 2)
    i0, i1, j0, j1 = is_ - nord, ie + nord, js - nord, je + nord + 1
    fy2 = put(fy2, i0, i1, j0, j1, S(g.del6_u, i0, i1, j0, j1) * (S(d2, i0, i1, j0 - 1, j1 - 1) - S(d2, i0, i1, j0, j1)))
    for n in range(1, nord + 1):
        nt = nord - n
        i0, i1, j0, j1 = is_ - nt - 1, ie + nt + 1, js - nt - 1, je + nt + 1
        d2 = put(torch.zeros_like(q), i0, i1, j0, j1,
                 (S(fx2, i0, i1, j0, j1) - S(fx2, i0 + 1, i1 + 1, j0, j1) + S(fy2, i0, i1, j0, j1) - S(fy2, i0, i1, j0 + 1, j1 + 1)) *
                 S(g.rarea, i0, i1, j0, j1))
        d2 = copy_corners(d2, npx, npy, 1)
        i0, i1, j0, j1 = is_ - nt, ie + nt + 1, js - nt, je + nt
        fx2 = put(torch.zeros_like(q), i0, i1, j0, j1, S(g.del6_v, i0, i1, j0, j1) * (S(d2, i0, i1, j0, j1) - S(d2, i0 - 1, i1 - 1, j0, j1)))
        d2 = copy_corners(d2, npx, npy, 2)
        i0, i1, j0, j1 = is_ - nt, ie + nt, js - nt, je + nt + 1
        fy2 = put(torch.zeros_like(q), i0, i1, j0, j1, S(g.del6_u, i0, i1, j0, j1) * (S(d2, i0, i1, j0, j1) - S(d2, i0, i1, j0 - 1, j1 - 1)))
    if mass is not None:
        damp2 = 0.5 * damp
        i0, i1, j0, j1 = is_, ie + 1, js, je
        fx = put(fx, i0, i1, j0, j1, S(fx, i0, i1, j0, j1) + damp2 * (S(mass, i0 - 1, i1 - 1, j0, j1) + S(mass, i0, i1, j0, j1)) * S(fx2, i0, i1, j0, j1))
        i0, i1, j0, j1 = is_, ie, js, je + 1
        fy = put(fy, i0, i1, j0, j1, S(fy, i0, i1, j0, j1) + damp2 * (S(mass, i0, i1, j0 - 1, j1 - 1) + S(mass, i0, i1, j0, j1)) * S(fy2, i0, i1, j0, j1))
    else:
        i0, i1, j0, j1 = is_, ie + 1, js, je
        fx = put(fx, i0, i1, j0, j1, S(fx, i0, i1, j0, j1) + S(fx2, i0, i1, j0, j1))
        i0, i1, j0, j1 = is_, ie, js, je + 1
        fy = put(fy, i0, i1, j0, j1, S(fy, i0, i1, j0, j1) + S(fy2, i0, i1, j0, j1))
    return fx, fy


def lev_select(vals_by_level, variants):
    """variants: dict key -> tensor [6,K,..]; vals_by_level: list (len K) of keys.  Returns the
    tensor that takes, for each level k, variants[vals_by_level[k]]."""
    keys = sorted(set(vals_by_level))
    if len(keys) == 1:
        return variants[keys[0]]
    out = None
    for key in keys:
        m = torch.tensor([v == key for v in vals_by_level], dtype=torch.bool).view(1, -1, 1, 1)
        out = variants[key] if out is None else torch.where(m, variants[key], out)
    return out


def fv_tp_2d_damp(q, crx, cry, hord, xfx, yfx, g, ra_x, ra_y, mfx=None, mfy=None, mass=None, nord=None, damp_c=None):
    """fv_tp_2d including the optional deln_flux (tp_core_nlm.F90:168-208).  nord / damp_c:
    per-level python lists (len K) or None."""
    fx, fy, q = fv_tp_2d(q, crx, cry, hord, xfx, yfx, g, ra_x, ra_y, mfx, mfy)
    if nord is None:
        return fx, fy, q
    K = q.shape[1]
    use_mass = (mfx is not None) and (mass is not None)
    if mfx is not None and mass is None:
        return fx, fy, q
    fxv, fyv, keys = {}, {}, []
    for k in range(K):
        if damp_c[k] > 1.e-4:
            keys.append((nord[k], damp_c[k]))
        else:
            keys.append(None)
    for key in set(keys):
        if key is None:
            fxv[key], fyv[key] = fx, fy
        else:
            damp = (key[1] * g.da_min) ** (key[0] + 1)
            fxv[key], fyv[key] = deln_flux(key[0], damp, q, fx, fy, g, mass if use_mass else None)
    skeys = [str(k) for k in keys]
    fx = lev_select(skeys, {str(k): v for k, v in fxv.items()})
    fy = lev_select(skeys, {str(k): v for k, v in fyv.items()})
    return fx, fy, q
