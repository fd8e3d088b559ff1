"""ORACLE (test infrastructure only).  d_sw, the D-grid shallow-water step, restated in torch
float64 from model/sw_core_nlm.F90:492-1545 (TL model_tlmadm/sw_core_tlm.F90:1047, AD
sw_core_adm.F90:1773/3269), with xtp_u :1970 / ytp_v :2312 (linear orders 1, 2),
del6_vt_flux :1547 and the divergence damping block :1264-1434.

Scope of the restatement (the configuration the TL/AD are run with):
grid_type 0, non-nested, non-stretched, inline_q = F, d_con = 0 (no dissipative heating),
do_f3d = F, no USE_COND / SW_DYNAMICS / OVERLOAD_R4.  Per-level parameters (sponge
layers, dyn_core_nlm.F90:579-625) are passed as python lists of length K.

parity: pinned by tests/test_ref_golden.py (the reference's D_SW_TLM and D_SW_FWD / D_SW_BWD executed on seeded inputs: values bit for
bit, tangents and adjoints to 6e-15) and tests/test_ref_tlm.py (XTP_U / YTP_V / DEL6_VT_FLUX _TLM).
"""
import torch
from .cubed_sphere import R, NG, copy_corners, fill_corners_bgrid, fill_corners_dgrid
from . import tp_core as tp
from .a2b_edge import a2b_ord4
from .sw_core import S, P, put, Z, sg, O, C1, C2, C3, P1, P2


# Test hook.  True: the damping of w carries its derivative, as in the reference's tangent-linear model (dyn_core_tlm.F90:856-858, 915-916 set
# nord_w_pert / damp_w_pert per level).  False reproduces the reference's REVERSE sweep, where dyn_core_adm.F90 only ever initialises those two
# switches to 0 (:338, :378, :1954, :1994), so D_SW_BWD runs without the adjoint of the w damping (tests/test_ref_golden.py).
W_DAMP_DERIVATIVE = True


class _Splice(torch.autograd.Function):
    """value of b (trajectory-scheme chain), derivative of a (perturbation-scheme chain): the split_hord / split_damp
    semantics of the TL/AD model (model_tlmadm/sw_core_tlm.F90:1664-1682, 2341-2366, 2436-2451)"""
    generate_vmap_rule = True

    @staticmethod
    def forward(a, b):
        return b.clone()

    @staticmethod
    def setup_context(ctx, inputs, output):
        pass

    @staticmethod
    def backward(ctx, g):
        return g, None

    @staticmethod
    def jvp(ctx, ta, tb):
        return ta


def splice(a, b):
    return _Splice.apply(a, b)


def _lv(vals, dtype=torch.float64):
    return torch.tensor(vals, dtype=dtype).view(1, -1, 1, 1)


def contravariant_winds(uc, vc, g, dt):
    """sw_core_nlm.F90:660-836: ut, vt from the C-grid winds incl. edge and corner 2x2 solves"""
    N, npx, npy = g.N, g.npx, g.npy
    is_, ie, js, je = 1, N, 1, N
    isd, ied, jsd, jed = g.isd, g.ied, g.jsd, g.jed
    sin1, _ = sg(g, 1); sin2, _ = sg(g, 2); sin3, _ = sg(g, 3); sin4, _ = sg(g, 4)
    cu, cv = g.cosa_u, g.cosa_v
    ut = Z(uc); vt = Z(uc)
    # interior
    for (j0, j1) in ((jsd, -1), (2, npy - 2), (npy + 1, jed)):
        i0, i1 = is_ - 1, ie + 2
        ut = put(ut, i0, i1, j0, j1, (S(uc, i0, i1, j0, j1) - 0.25 * S(cu, i0, i1, j0, j1) *
                                      (S(vc, i0 - 1, i1 - 1, j0, j1) + S(vc, i0, i1, j0, j1) + S(vc, i0 - 1, i1 - 1, j0 + 1, j1 + 1) + S(vc, i0, i1, j0 + 1, j1 + 1))) *
                 S(g.rsin_u, i0, i1, j0, j1))
    for (j0, j1) in ((js - 1, 0), (2, npy - 1), (npy + 1, je + 2)):
        i0, i1 = isd, ied
        vt = put(vt, i0, i1, j0, j1, (S(vc, i0, i1, j0, j1) - 0.25 * S(cv, i0, i1, j0, j1) *
                                      (S(uc, i0, i1, j0 - 1, j1 - 1) + S(uc, i0 + 1, i1 + 1, j0 - 1, j1 - 1) + S(uc, i0, i1, j0, j1) + S(uc, i0 + 1, i1 + 1, j0, j1))) *
                 S(g.rsin_v, i0, i1, j0, j1))
    # west / east edges
    for i in (1, npx):
        ucv = S(uc, i, i, jsd, jed)
        ut = put(ut, i, i, jsd, jed, torch.where(ucv * dt > 0., ucv / S(sin3, i - 1, i - 1, jsd, jed), ucv / S(sin1, i, i, jsd, jed)))
    j0, j1 = 3, npy - 2
    for i in (0, 1, npx - 1, npx):
        vt = put(vt, i, i, j0, j1, S(vc, i, i, j0, j1) - 0.25 * S(cv, i, i, j0, j1) *
                 (S(ut, i, i, j0 - 1, j1 - 1) + S(ut, i + 1, i + 1, j0 - 1, j1 - 1) + S(ut, i, i, j0, j1) + S(ut, i + 1, i + 1, j0, j1)))
    # south / north edges
    for j in (1, npy):
        vcv = S(vc, isd, ied, j, j)
        vt = put(vt, isd, ied, j, j, torch.where(vcv * dt > 0., vcv / S(sin4, isd, ied, j - 1, j - 1), vcv / S(sin2, isd, ied, j, j)))
    i0, i1 = 3, npx - 2
    for j in (0, 1, npy - 1, npy):
        ut = put(ut, i0, i1, j, j, S(uc, i0, i1, j, j) - 0.25 * S(cu, i0, i1, j, j) *
                 (S(vt, i0 - 1, i1 - 1, j, j) + S(vt, i0, i1, j, j) + S(vt, i0 - 1, i1 - 1, j + 1, j + 1) + S(vt, i0, i1, j + 1, j + 1)))
    # corners (2x2 solves).  All right-hand sides use values already final
    ut = ut.clone(); vt = vt.clone()
    U, V, UC, VC = (lambda i, j: P(ut, i, j)), (lambda i, j: P(vt, i, j)), (lambda i, j: P(uc, i, j)), (lambda i, j: P(vc, i, j))
    CU, CV = (lambda i, j: P(cu, i, j)), (lambda i, j: P(cv, i, j))
    def setu(i, j, v): ut[..., j + O, i + O] = v
    def setv(i, j, v): vt[..., j + O, i + O] = v
    # sw
    d = 1. / (1. - 0.0625 * CU(2, 0) * CV(1, 0))
    setu(2, 0, (UC(2, 0) - 0.25 * CU(2, 0) * (V(1, 1) + V(2, 1) + V(2, 0) + VC(1, 0) - 0.25 * CV(1, 0) * (U(1, 0) + U(1, -1) + U(2, -1)))) * d)
    d = 1. / (1. - 0.0625 * CU(0, 1) * CV(0, 2))
    setv(0, 2, (VC(0, 2) - 0.25 * CV(0, 2) * (U(1, 1) + U(1, 2) + U(0, 2) + UC(0, 1) - 0.25 * CU(0, 1) * (V(0, 1) + V(-1, 1) + V(-1, 2)))) * d)
    d = 1. / (1. - 0.0625 * CU(2, 1) * CV(1, 2))
    u21 = (UC(2, 1) - 0.25 * CU(2, 1) * (V(1, 1) + V(2, 1) + V(2, 2) + VC(1, 2) - 0.25 * CV(1, 2) * (U(1, 1) + U(1, 2) + U(2, 2)))) * d
    v12 = (VC(1, 2) - 0.25 * CV(1, 2) * (U(1, 1) + U(1, 2) + U(2, 2) + UC(2, 1) - 0.25 * CU(2, 1) * (V(1, 1) + V(2, 1) + V(2, 2)))) * d
    setu(2, 1, u21); setv(1, 2, v12)
    # se
    n = npx
    d = 1. / (1. - 0.0625 * CU(n - 1, 0) * CV(n - 1, 0))
    setu(n - 1, 0, (UC(n - 1, 0) - 0.25 * CU(n - 1, 0) * (V(n - 1, 1) + V(n - 2, 1) + V(n - 2, 0) + VC(n - 1, 0) - 0.25 * CV(n - 1, 0) * (U(n, 0) + U(n, -1) + U(n - 1, -1)))) * d)
    d = 1. / (1. - 0.0625 * CU(n + 1, 1) * CV(n, 2))
    setv(n, 2, (VC(n, 2) - 0.25 * CV(n, 2) * (U(n, 1) + U(n, 2) + U(n + 1, 2) + UC(n + 1, 1) - 0.25 * CU(n + 1, 1) * (V(n, 1) + V(n + 1, 1) + V(n + 1, 2)))) * d)
    d = 1. / (1. - 0.0625 * CU(n - 1, 1) * CV(n - 1, 2))
    u_ = (UC(n - 1, 1) - 0.25 * CU(n - 1, 1) * (V(n - 1, 1) + V(n - 2, 1) + V(n - 2, 2) + VC(n - 1, 2) - 0.25 * CV(n - 1, 2) * (U(n, 1) + U(n, 2) + U(n - 1, 2)))) * d
    v_ = (VC(n - 1, 2) - 0.25 * CV(n - 1, 2) * (U(n, 1) + U(n, 2) + U(n - 1, 2) + UC(n - 1, 1) - 0.25 * CU(n - 1, 1) * (V(n - 1, 1) + V(n - 2, 1) + V(n - 2, 2)))) * d
    setu(n - 1, 1, u_); setv(n - 1, 2, v_)
    # ne
    m = npy
    d = 1. / (1. - 0.0625 * CU(n - 1, m) * CV(n - 1, m + 1))
    setu(n - 1, m, (UC(n - 1, m) - 0.25 * CU(n - 1, m) * (V(n - 1, m) + V(n - 2, m) + V(n - 2, m + 1) + VC(n - 1, m + 1) - 0.25 * CV(n - 1, m + 1) * (U(n, m) + U(n, m + 1) + U(n - 1, m + 1)))) * d)
    d = 1. / (1. - 0.0625 * CU(n + 1, m - 1) * CV(n, m - 1))
    setv(n, m - 1, (VC(n, m - 1) - 0.25 * CV(n, m - 1) * (U(n, m - 1) + U(n, m - 2) + U(n + 1, m - 2) + UC(n + 1, m - 1) - 0.25 * CU(n + 1, m - 1) * (V(n, m) + V(n + 1, m) + V(n + 1, m - 1)))) * d)
    d = 1. / (1. - 0.0625 * CU(n - 1, m - 1) * CV(n - 1, m - 1))
    u_ = (UC(n - 1, m - 1) - 0.25 * CU(n - 1, m - 1) * (V(n - 1, m) + V(n - 2, m) + V(n - 2, m - 1) + VC(n - 1, m - 1) - 0.25 * CV(n - 1, m - 1) * (U(n, m - 1) + U(n, m - 2) + U(n - 1, m - 2)))) * d
    v_ = (VC(n - 1, m - 1) - 0.25 * CV(n - 1, m - 1) * (U(n, m - 1) + U(n, m - 2) + U(n - 1, m - 2) + UC(n - 1, m - 1) - 0.25 * CU(n - 1, m - 1) * (V(n - 1, m) + V(n - 2, m) + V(n - 2, m - 1)))) * d
    setu(n - 1, m - 1, u_); setv(n - 1, m - 1, v_)
    # nw
    d = 1. / (1. - 0.0625 * CU(2, m) * CV(1, m + 1))
    setu(2, m, (UC(2, m) - 0.25 * CU(2, m) * (V(1, m) + V(2, m) + V(2, m + 1) + VC(1, m + 1) - 0.25 * CV(1, m + 1) * (U(1, m) + U(1, m + 1) + U(2, m + 1)))) * d)
    d = 1. / (1. - 0.0625 * CU(0, m - 1) * CV(0, m - 1))
    setv(0, m - 1, (VC(0, m - 1) - 0.25 * CV(0, m - 1) * (U(1, m - 1) + U(1, m - 2) + U(0, m - 2) + UC(0, m - 1) - 0.25 * CU(0, m - 1) * (V(0, m) + V(-1, m) + V(-1, m - 1)))) * d)
    d = 1. / (1. - 0.0625 * CU(2, m - 1) * CV(1, m - 1))
    u_ = (UC(2, m - 1) - 0.25 * CU(2, m - 1) * (V(1, m) + V(2, m) + V(2, m - 1) + VC(1, m - 1) - 0.25 * CV(1, m - 1) * (U(1, m - 1) + U(1, m - 2) + U(2, m - 2)))) * d
    v_ = (VC(1, m - 1) - 0.25 * CV(1, m - 1) * (U(1, m - 1) + U(1, m - 2) + U(2, m - 2) + UC(2, m - 1) - 0.25 * CU(2, m - 1) * (V(1, m) + V(2, m) + V(2, m - 1)))) * d
    setu(2, m - 1, u_); setv(1, m - 1, v_)
    return ut, vt


def courant_fluxes(ut, vt, g, dt):
    """sw_core_nlm.F90:853-907: crx, cry, xfx, yfx, ra_x, ra_y"""
    N = g.N
    is_, ie, js, je = 1, N, 1, N
    isd, ied, jsd, jed = g.isd, g.ied, g.jsd, g.jed
    sin1, _ = sg(g, 1); sin2, _ = sg(g, 2); sin3, _ = sg(g, 3); sin4, _ = sg(g, 4)
    i0, i1, j0, j1 = is_, ie + 1, jsd, jed
    xf = dt * S(ut, i0, i1, j0, j1)
    pos = xf > 0.
    crx = put(Z(ut), i0, i1, j0, j1, torch.where(pos, xf * S(g.rdxa, i0 - 1, i1 - 1, j0, j1), xf * S(g.rdxa, i0, i1, j0, j1)))
    xfx = put(Z(ut), i0, i1, j0, j1, S(g.dy, i0, i1, j0, j1) * xf * torch.where(pos, S(sin3, i0 - 1, i1 - 1, j0, j1), S(sin1, i0, i1, j0, j1)))
    i0, i1, j0, j1 = isd, ied, js, je + 1
    yf = dt * S(vt, i0, i1, j0, j1)
    pos = yf > 0.
    cry = put(Z(ut), i0, i1, j0, j1, torch.where(pos, yf * S(g.rdya, i0, i1, j0 - 1, j1 - 1), yf * S(g.rdya, i0, i1, j0, j1)))
    yfx = put(Z(ut), i0, i1, j0, j1, S(g.dx, i0, i1, j0, j1) * yf * torch.where(pos, S(sin4, i0, i1, j0 - 1, j1 - 1), S(sin2, i0, i1, j0, j1)))
    ra_x = put(Z(ut), is_, ie, jsd, jed, S(g.area, is_, ie, jsd, jed) + (S(xfx, is_, ie, jsd, jed) - S(xfx, is_ + 1, ie + 1, jsd, jed)))
    ra_y = put(Z(ut), isd, ied, js, je, S(g.area, isd, ied, js, je) + (S(yfx, isd, ied, js, je) - S(yfx, isd, ied, js + 1, je + 1)))
    return crx, cry, xfx, yfx, ra_x, ra_y


def xtp_u(c, u, g, iord):
    """sw_core_nlm.F90:1970-2309 (iord 1, 2) and sw_core_tlm.F90:7332-7356 (iord 333).  flux(is:ie+1, js:je+1); u is the D-grid u.
    iord: int or per-level list."""
    N, npx, npy = g.N, g.npx, g.npy
    is_, ie, js, je = 1, N, 1, N
    j0, j1 = js, je + 1
    dx, rdx = g.dx, g.rdx
    cc = S(c, is_, ie + 1, j0, j1)
    um = S(u, is_ - 1, ie, j0, j1); up_ = S(u, is_, ie + 1, j0, j1)
    f1 = torch.where(cc > 0., um, up_)
    # bl, br for cells 0..npx
    lo, hi = 0, npx
    def U(a, b): return S(u, a, b, j0, j1)
    al = P1 * (U(lo - 1, hi) + U(lo, hi + 1)) + P2 * (U(lo - 2, hi - 1) + U(lo + 1, hi + 2))     # al(i), i = lo..hi+1
    ucell = U(lo, hi)
    bl = al[..., :-1] - ucell
    br = al[..., 1:] - ucell
    bl = list(torch.unbind(bl, -1)); br = list(torch.unbind(br, -1))
    def u1(i): return u[..., R(j0, j1), i + O]
    def d1(i): return dx[..., R(j0, j1), i + O]
    def two_sided(i):   # value at the face between cells i-1 and i
        return 0.5 * (((2. * d1(i - 1) + d1(i - 2)) * u1(i - 1) - d1(i - 1) * u1(i - 2)) / (d1(i - 1) + d1(i - 2)) +
                      ((2. * d1(i) + d1(i + 1)) * u1(i) - d1(i) * u1(i + 1)) / (d1(i) + d1(i + 1)))
    jj = torch.arange(j0, j1 + 1)
    edge_row = ((jj == 1) | (jj == npy)).view(1, 1, -1)
    def zrow(v): return torch.where(edge_row, torch.zeros_like(v), v)
    # west
    xt = C3 * u1(1) + C2 * u1(2) + C1 * u1(3)
    br[1] = xt - u1(1); bl[2] = xt - u1(2); br[2] = al[..., 3] - u1(2)
    bl[0] = C1 * u1(-2) + C2 * u1(-1) + C3 * u1(0) - u1(0)
    xt = two_sided(1)
    br[0] = xt - u1(0); bl[1] = xt - u1(1)
    bl[0] = zrow(bl[0]); br[0] = zrow(br[0]); bl[1] = zrow(bl[1]); br[1] = zrow(br[1])
    # east
    bl[npx - 2] = al[..., npx - 2] - u1(npx - 2)
    xt = C1 * u1(npx - 3) + C2 * u1(npx - 2) + C3 * u1(npx - 1)
    br[npx - 2] = xt - u1(npx - 2); bl[npx - 1] = xt - u1(npx - 1)
    xt = two_sided(npx)
    br[npx - 1] = xt - u1(npx - 1); bl[npx] = xt - u1(npx)
    br[npx] = C3 * u1(npx) + C2 * u1(npx + 1) + C1 * u1(npx + 2) - u1(npx)
    bl[npx - 1] = zrow(bl[npx - 1]); br[npx - 1] = zrow(br[npx - 1]); bl[npx] = zrow(bl[npx]); br[npx] = zrow(br[npx])
    bl = torch.stack(bl, -1); br = torch.stack(br, -1)
    b0 = bl + br
    cfl_p = cc * S(rdx, is_ - 1, ie, j0, j1)
    cfl_n = cc * S(rdx, is_, ie + 1, j0, j1)
    f_pos = um + (1. - cfl_p) * (br[..., is_ - 1: ie + 1] - cfl_p * b0[..., is_ - 1: ie + 1])
    f_neg = up_ + (1. + cfl_n) * (bl[..., is_: ie + 2] + cfl_n * b0[..., is_: ie + 2])
    f2 = torch.where(cc > 0., f_pos, f_neg)
    table = {1: f1, 2: f2}
    for o in set([iord] if isinstance(iord, int) else iord):
        if 8 <= o <= 13:
            # monotone schemes of the nonlinear model (sw_core_nlm.F90:2166-2306): iord 8, 9, 10, else unlimited
            near_zero = 1.e-9
            def DM(a, b):
                xt = 0.25 * (U(a + 1, b + 1) - U(a - 1, b - 1)); c0 = U(a, b)
                return tp._sign(tp._min3(xt.abs(), tp._max3(U(a - 1, b - 1), c0, U(a + 1, b + 1)) - c0, c0 - tp._min3(U(a - 1, b - 1), c0, U(a + 1, b + 1))), xt)
            DQ = lambda a, b: U(a + 1, b + 1) - U(a, b)
            alm = 0.5 * (U(lo - 1, hi) + U(lo, hi + 1)) + tp.R3 * (DM(lo - 1, hi) - DM(lo, hi + 1))
            u0 = U(lo, hi)
            mbl, mbr = alm[..., :-1] - u0, alm[..., 1:] - u0
            z = torch.zeros_like(mbl)
            if o == 8:
                xt = 2. * DM(lo, hi)
                mbl, mbr = -tp._sign(torch.minimum(xt.abs(), mbl.abs()), xt), tp._sign(torch.minimum(xt.abs(), mbr.abs()), xt)
            elif o in (9, 10):
                pmp_1 = -2. * DQ(lo, hi); lac_1 = pmp_1 + 1.5 * DQ(lo + 1, hi + 1)
                bl_l = torch.minimum(tp._max3(z, pmp_1, lac_1), torch.maximum(mbl, tp._min3(z, pmp_1, lac_1)))
                pmp_2 = 2. * DQ(lo - 1, hi - 1); lac_2 = pmp_2 - 1.5 * DQ(lo - 2, hi - 2)
                br_l = torch.minimum(tp._max3(z, pmp_2, lac_2), torch.maximum(mbr, tp._min3(z, pmp_2, lac_2)))
                if o == 9:
                    mbl, mbr = bl_l, br_l
                else:
                    small = DM(lo, hi).abs() < near_zero
                    flat = (DM(lo - 1, hi - 1).abs() + DM(lo + 1, hi + 1).abs()) < near_zero
                    lim = (3. * (mbl + mbr)).abs() > (mbl - mbr).abs()
                    mbl, mbr = (torch.where(small, torch.where(flat, z, mbl), torch.where(lim, bl_l, mbl)),
                                torch.where(small, torch.where(flat, z, mbr), torch.where(lim, br_l, mbr)))
            mbl = list(torch.unbind(mbl, -1)); mbr = list(torch.unbind(mbr, -1))
            dmv = list(torch.unbind(DM(-1, npx + 1), -1)); dm = lambda i: dmv[i + 1]
            def x0(i):      # x0L + x0R at the cube edge between cells i-1 and i (no limiter here)
                return (0.5 * ((2. * d1(i - 1) + d1(i - 2)) * u1(i - 1) - d1(i - 1) * u1(i - 2)) / (d1(i - 1) + d1(i - 2)) +
                        0.5 * ((2. * d1(i) + d1(i + 1)) * u1(i) - d1(i) * u1(i + 1)) / (d1(i) + d1(i + 1)))
            mbr[2] = alm[..., 3] - u1(2)
            xt = tp.S15 * u1(1) + tp.S11 * u1(2) - tp.S14 * dm(2)
            mbl[2] = xt - u1(2); mbr[1] = xt - u1(1)
            mbl[0] = tp.S14 * dm(-1) - tp.S11 * (u1(0) - u1(-1))
            xt = x0(1)
            mbr[0] = xt - u1(0); mbl[1] = xt - u1(1)
            mbl[npx - 2] = alm[..., npx - 2] - u1(npx - 2)
            xt = tp.S15 * u1(npx - 1) + tp.S11 * u1(npx - 2) + tp.S14 * dm(npx - 2)
            mbr[npx - 2] = xt - u1(npx - 2); mbl[npx - 1] = xt - u1(npx - 1)
            mbr[npx] = tp.S11 * (u1(npx + 1) - u1(npx)) - tp.S14 * dm(npx + 1)
            xt = x0(npx)
            mbr[npx - 1] = xt - u1(npx - 1); mbl[npx] = xt - u1(npx)
            for i in (0, 1, npx - 1, npx):
                mbl[i] = zrow(mbl[i]); mbr[i] = zrow(mbr[i])
            for i in (2, npx - 2):
                mbl[i], mbr[i] = tp.pert_ppm(u1(i), mbl[i], mbr[i], -1)
            mbl = torch.stack(mbl, -1); mbr = torch.stack(mbr, -1)
            fp = um + (1. - cfl_p) * (mbr[..., is_ - 1: ie + 1] - cfl_p * (mbl[..., is_ - 1: ie + 1] + mbr[..., is_ - 1: ie + 1]))
            fn_ = up_ + (1. + cfl_n) * (mbl[..., is_: ie + 2] + cfl_n * (mbl[..., is_: ie + 2] + mbr[..., is_: ie + 2]))
            table[o] = torch.where(cc > 0., fp, fn_)
        if 3 <= o <= 7:        # smoothness-switch schemes (sw_core_nlm.F90:2082-2160) on the bl / br of the linear scheme
            sl = lambda a, d: a[..., is_ + d: ie + 2 + d]
            blm, brm, b0m, blp, brp, b0p = sl(bl, -1), sl(br, -1), sl(b0, -1), sl(bl, 0), sl(br, 0), sl(b0, 0)
            z = torch.zeros_like(cc)
            if o in (3, 4):
                s5 = b0.abs() < (bl - br).abs(); s6 = 3. * b0.abs() < (bl - br).abs()
                s5m, s5p, s6m, s6p = sl(s5, -1), sl(s5, 0), sl(s6, -1), sl(s6, 0)
                if o == 3:
                    f0p = torch.where(s6m | s5p, brm - cfl_p * b0m, torch.where(s5m, tp._sign(torch.minimum(blm.abs(), brm.abs()), brm), z))
                    f0n = torch.where(s6p | s5m, blp + cfl_n * b0p, torch.where(s5p, tp._sign(torch.minimum(blp.abs(), brp.abs()), blp), z))
                    table[o] = torch.where(cc > 0., um + (1. - cfl_p) * f0p, up_ + (1. + cfl_n) * f0n)
                else:
                    table[o] = torch.where(cc > 0., um + torch.where(s6m | s5p, (1. - cfl_p) * (brm - cfl_p * b0m), z),
                                           up_ + torch.where(s6p | s5m, (1. + cfl_n) * (blp + cfl_n * b0p), z))
            else:
                s5 = (bl * br < 0.) if o == 5 else ((3. * b0).abs() < (bl - br).abs())
                on = sl(s5, -1) | sl(s5, 0)
                f0 = torch.where(cc > 0., (1. - cfl_p) * (brm - cfl_p * b0m), (1. + cfl_n) * (blp + cfl_n * b0p))
                table[o] = f1 + torch.where(on, f0, z)
    # iord = 333 (sw_core_tlm.F90:7332-7356): third-order linear, Courant number c * rdx of the upwind cell
    umm = S(u, is_ - 2, ie - 1, j0, j1); upp = S(u, is_ + 1, ie + 2, j0, j1)
    f3 = torch.where(cc > 0., (2.0 * up_ + 5.0 * um - umm) / 6.0 - 0.5 * cfl_p * (up_ - um) + cfl_p * cfl_p / 6.0 * (up_ - 2.0 * um + umm),
                     (2.0 * um + 5.0 * up_ - upp) / 6.0 - 0.5 * cfl_n * (up_ - um) + cfl_n * cfl_n / 6.0 * (upp - 2.0 * up_ + um))
    table[333] = f3
    return put(Z(u), is_, ie + 1, j0, j1, tp.select_ord(iord, table))


def ytp_v(c, v, g, jord):
    """sw_core_nlm.F90:2312-2738 is the transpose of xtp_u (dy for dx, npy for npx)."""
    class GT:
        pass
    gt = GT()
    gt.N = g.N; gt.npx = g.npy; gt.npy = g.npx
    gt.dx = g.dy.transpose(-1, -2); gt.rdx = g.rdy.transpose(-1, -2)
    return xtp_u(c.transpose(-1, -2), v.transpose(-1, -2), gt, jord).transpose(-1, -2)


def del6_vt_flux(nord, damp, q, g):
    """sw_core_nlm.F90:1547-1659: returns fx2, fy2 (and d2).  nord python int."""
    N, npx, npy = g.N, g.npx, g.npy
    is_, ie, js, je = 1, N, 1, N
    i1_, i2_, j1_, j2_ = is_ - 1 - nord, ie + 1 + nord, js - 1 - nord, je + 1 + nord
    d2 = put(Z(q), i1_, i2_, j1_, j2_, damp * S(q, i1_, i2_, j1_, j2_))
    if nord > 0:
        d2 = copy_corners(d2, npx, npy, 1)
    i0, i1, j0, j1 = is_ - nord, ie + nord + 1, js - nord, je + nord
    fx2 = put(Z(q), i0, i1, j0, j1, S(g.del6_v, i0, i1, j0, j1) * (S(d2, i0 - 1, i1 - 1, j0, j1) - S(d2, i0, i1, j0, j1)))
    if nord > 0:
        d2 = copy_corners(d2, npx, npy, 2)
    i0, i1, j0, j1 = is_ - nord, ie + nord, js - nord, je + nord + 1
    fy2 = put(Z(q), i0, i1, j0, j1, S(g.del6_u, i0, i1, j0, j1) * (S(d2, i0, i1, j0 - 1, j1 - 1) - S(d2, i0, i1, j0, j1)))
    for n in range(1, nord + 1):
        nt = nord - n
        i0, i1, j0, j1 = is_ - nt - 1, ie + nt + 1, js - nt - 1, je + nt + 1
        d2 = put(Z(q), i0, i1, j0, j1, ((S(fx2, i0, i1, j0, j1) - S(fx2, i0 + 1, i1 + 1, j0, j1)) + (S(fy2, i0, i1, j0, j1) - S(fy2, i0, i1, j0 + 1, j1 + 1))) *
                 S(g.rarea, i0, i1, j0, j1))
        d2 = copy_corners(d2, npx, npy, 1)
        i0, i1, j0, j1 = is_ - nt, ie + nt + 1, js - nt, je + nt
        fx2 = put(Z(q), i0, i1, j0, j1, S(g.del6_v, i0, i1, j0, j1) * (S(d2, i0, i1, j0, j1) - S(d2, i0 - 1, i1 - 1, j0, j1)))
        d2 = copy_corners(d2, npx, npy, 2)
        i0, i1, j0, j1 = is_ - nt, ie + nt, js - nt, je + nt + 1
        fy2 = put(Z(q), i0, i1, j0, j1, S(g.del6_u, i0, i1, j0, j1) * (S(d2, i0, i1, j0, j1) - S(d2, i0, i1, j0 - 1, j1 - 1)))
    return fx2, fy2


def del6_by_level(nord_l, damp_l, q, g):
    """per-level (nord, damp) lists -> fx2, fy2"""
    keys = [(n, d) for n, d in zip(nord_l, damp_l)]
    fxv, fyv = {}, {}
    for key in set(keys):
        fxv[str(key)], fyv[str(key)] = del6_vt_flux(key[0], key[1], q, g)
    sk = [str(k) for k in keys]
    return tp.lev_select(sk, fxv), tp.lev_select(sk, fyv)


def d_sw(delp, pt, u, v, w, uc, vc, ua, va, divg_d, g, dt, prm, pp=None):
    """prm: dict with per-level lists hord_mt, hord_vt, hord_tm, hord_dp, nord, nord_v, nord_w,
    nord_t, d2_bg, damp_v, damp_w, damp_t and scalars dddmp, d4_bg, hydrostatic.
    pp: the perturbation-side switches (same keys + split_damp) of D_SW_TLM (model_tlmadm/sw_core_tlm.F90:1047); where they
    differ from prm the operator is evaluated with both and spliced (value from prm, derivative from pp).
    Returns dict(delp, pt, u, v, w, fx, fy, crx, cry, xfx, yfx)."""
    if pp is None:
        pp = prm
    split_damp = bool(pp.get("split_damp", False))

    def tp_site(same, fn):
        """fn(p) -> (fx, fy, q) with the switches p"""
        if same:
            return fn(prm)
        a = fn(pp); b = fn(prm)
        return splice(a[0], b[0]), splice(a[1], b[1]), b[2]
    N, npx, npy = g.N, g.npx, g.npy
    is_, ie, js, je = 1, N, 1, N
    isd, ied, jsd, jed = g.isd, g.ied, g.jsd, g.jed
    K = delp.shape[1]
    hydro = prm["hydrostatic"]
    sin1, _ = sg(g, 1); sin2, _ = sg(g, 2); sin3, _ = sg(g, 3); sin4, _ = sg(g, 4)
    rarea = g.rarea

    ut, vt = contravariant_winds(uc, vc, g, dt)
    crx, cry, xfx, yfx, ra_x, ra_y = courant_fluxes(ut, vt, g, dt)

    def div(fx_, fy_, i0=is_, i1=ie, j0=js, j1=je):
        return ((S(fx_, i0, i1, j0, j1) - S(fx_, i0 + 1, i1 + 1, j0, j1)) + (S(fy_, i0, i1, j0, j1) - S(fy_, i0, i1, j0 + 1, j1 + 1))) * S(rarea, i0, i1, j0, j1)

    fx, fy, delp_c = tp_site(prm["hord_dp"] == pp["hord_dp"] and not split_damp,
                             lambda p: tp.fv_tp_2d_damp(delp, crx, cry, p["hord_dp"], xfx, yfx, g, ra_x, ra_y, nord=p["nord_v"], damp_c=p["damp_v"]))

    w_new = w
    if not hydro:
        dmpw = prm["damp_w"]; nordw = prm["nord_w"]
        damp4 = [(dmpw[k] * g.da_min_c) ** (nordw[k] + 1) if dmpw[k] > 1.e-5 else 0.0 for k in range(K)]
        fx2, fy2 = del6_by_level(nordw, damp4, w, g)
        if not W_DAMP_DERIVATIVE:
            fx2, fy2 = fx2.detach(), fy2.detach()
        on = _lv([1.0 if dmpw[k] > 1.e-5 else 0.0 for k in range(K)])
        dw = div(fx2, fy2) * on
        gx, gy, _ = tp_site(prm["hord_vt"] == pp["hord_vt"], lambda p: tp.fv_tp_2d(w, crx, cry, p["hord_vt"], xfx, yfx, g, ra_x, ra_y, mfx=fx, mfy=fy))
        w_new = put(w, is_, ie, js, je, S(delp, is_, ie, js, je) * S(w, is_, ie, js, je) + div(gx, gy))

    gx, gy, pt_c = tp_site(prm["hord_tm"] == pp["hord_tm"] and not split_damp,
                           lambda p: tp.fv_tp_2d_damp(pt, crx, cry, p["hord_tm"], xfx, yfx, g, ra_x, ra_y, mfx=fx, mfy=fy, mass=delp,
                                                      nord=p["nord_t"], damp_c=p["damp_t"]))
    ptd = S(pt, is_, ie, js, je) * S(delp, is_, ie, js, je) + div(gx, gy)
    delp_new_v = S(delp, is_, ie, js, je) + div(fx, fy)
    delp_new = put(delp, is_, ie, js, je, delp_new_v)
    pt_new = put(pt, is_, ie, js, je, ptd / delp_new_v)

    # ---- kinetic energy (:1051-1202)
    dt4, dt5, dt6 = 0.25 * dt, 0.5 * dt, dt / 6.
    i0, i1, j0, j1 = 2, npx - 1, 2, npy - 1
    vb = Z(u)
    vb = put(vb, is_, ie + 1, 1, 1, dt5 * (S(vt, is_ - 1, ie, 1, 1) + S(vt, is_, ie + 1, 1, 1)))
    vb = put(vb, i0, i1, j0, j1, dt5 * ((S(vc, i0 - 1, i1 - 1, j0, j1) + S(vc, i0, i1, j0, j1)) - (S(uc, i0, i1, j0 - 1, j1 - 1) + S(uc, i0, i1, j0, j1)) * S(g.cosa, i0, i1, j0, j1)) * S(g.rsina, i0, i1, j0, j1))
    for i in (1, npx):
        vb = put(vb, i, i, j0, j1, dt4 * (-S(vt, i - 2, i - 2, j0, j1) + 3. * (S(vt, i - 1, i - 1, j0, j1) + S(vt, i, i, j0, j1)) - S(vt, i + 1, i + 1, j0, j1)))
    vb = put(vb, is_, ie + 1, npy, npy, dt5 * (S(vt, is_ - 1, ie, npy, npy) + S(vt, is_, ie + 1, npy, npy)))
    same_mt = prm["hord_mt"] == pp["hord_mt"]
    ub = ytp_v(vb, v, g, prm["hord_mt"]) if same_mt else splice(ytp_v(vb, v, g, pp["hord_mt"]), ytp_v(vb, v, g, prm["hord_mt"]))
    ke = vb * ub
    ub2 = Z(u)
    ub2 = put(ub2, 1, 1, js, je + 1, dt5 * (S(ut, 1, 1, js - 1, je) + S(ut, 1, 1, js, je + 1)))
    ub2 = put(ub2, i0, i1, j0, j1, dt5 * ((S(uc, i0, i1, j0 - 1, j1 - 1) + S(uc, i0, i1, j0, j1)) - (S(vc, i0 - 1, i1 - 1, j0, j1) + S(vc, i0, i1, j0, j1)) * S(g.cosa, i0, i1, j0, j1)) * S(g.rsina, i0, i1, j0, j1))
    for j in (1, npy):
        ub2 = put(ub2, i0, i1, j, j, dt4 * (-S(ut, i0, i1, j - 2, j - 2) + 3. * (S(ut, i0, i1, j - 1, j - 1) + S(ut, i0, i1, j, j)) - S(ut, i0, i1, j + 1, j + 1)))
    ub2 = put(ub2, npx, npx, js, je + 1, dt5 * (S(ut, npx, npx, js - 1, je) + S(ut, npx, npx, js, je + 1)))
    vb2 = xtp_u(ub2, u, g, prm["hord_mt"]) if same_mt else splice(xtp_u(ub2, u, g, pp["hord_mt"]), xtp_u(ub2, u, g, prm["hord_mt"]))
    ke = 0.5 * (ke + ub2 * vb2)
    ke = ke.clone()
    U_, V_, UT, VT = (lambda i, j: P(u, i, j)), (lambda i, j: P(v, i, j)), (lambda i, j: P(ut, i, j)), (lambda i, j: P(vt, i, j))
    ke[..., 1 + O, 1 + O] = dt6 * ((UT(1, 1) + UT(1, 0)) * U_(1, 1) + (VT(1, 1) + VT(0, 1)) * V_(1, 1) + (UT(1, 1) + VT(1, 1)) * U_(0, 1))
    i = npx
    ke[..., 1 + O, i + O] = dt6 * ((UT(i, 1) + UT(i, 0)) * U_(i - 1, 1) + (VT(i, 1) + VT(i - 1, 1)) * V_(i, 1) + (UT(i, 1) - VT(i - 1, 1)) * U_(i, 1))
    j = npy
    ke[..., j + O, i + O] = dt6 * ((UT(i, j) + UT(i, j - 1)) * U_(i - 1, j) + (VT(i, j) + VT(i - 1, j)) * V_(i, j - 1) + (UT(i, j - 1) + VT(i - 1, j)) * U_(i, j))
    ke[..., j + O, 1 + O] = dt6 * ((UT(1, j) + UT(1, j - 1)) * U_(1, j) + (VT(1, j) + VT(0, j)) * V_(1, j - 1) + (UT(1, j - 1) - VT(1, j)) * U_(0, j))

    # ---- relative vorticity (:1204-1221)
    vt2 = put(Z(u), isd, ied, jsd, jed + 1, S(u, isd, ied, jsd, jed + 1) * S(g.dx, isd, ied, jsd, jed + 1))
    ut2 = put(Z(u), isd, ied + 1, jsd, jed, S(v, isd, ied + 1, jsd, jed) * S(g.dy, isd, ied + 1, jsd, jed))
    wk = put(Z(u), isd, ied, jsd, jed, S(rarea, isd, ied, jsd, jed) * ((S(vt2, isd, ied, jsd, jed) - S(vt2, isd, ied, jsd + 1, jed + 1)) +
                                                                    (S(ut2, isd + 1, ied + 1, jsd, jed) - S(ut2, isd, ied, jsd, jed))))
    if not hydro:
        wv = S(w_new, is_, ie, js, je) / delp_new_v + dw
        w_new = put(w_new, is_, ie, js, je, wv)

    # ---- divergence damping (:1264-1434), per level nord = 0 or > 0
    C = (is_, ie + 1, js, je + 1)
    def div_damp(p):
        nord_l = p["nord"]; d2bg = _lv(p["d2_bg"]); dddmp = p["dddmp"]; d4_bg = p["d4_bg"]
        da_min_c = g.da_min_c
        C = (is_, ie + 1, js, je + 1)
        ke_damp = {}
        divg = {}          # what d_sw leaves in delpc (the caller's vt, dyn_core_nlm.F90:656): its own divergence (:1340) or divg_d (:1353)
        if any(n == 0 for n in nord_l):
            # nord == 0 :  del-2 on the divergence computed from the D-grid winds
            ptc = Z(u)
            j0, j1 = 2, npy - 1
            ptc = put(ptc, is_ - 1, ie + 1, j0, j1, (S(u, is_ - 1, ie + 1, j0, j1) - 0.5 * (S(va, is_ - 1, ie + 1, j0 - 1, j1 - 1) + S(va, is_ - 1, ie + 1, j0, j1)) * S(g.cosa_v, is_ - 1, ie + 1, j0, j1)) *
                      S(g.dyc, is_ - 1, ie + 1, j0, j1) * S(g.sina_v, is_ - 1, ie + 1, j0, j1))
            for j in (1, npy):
                vcv = S(vc, is_ - 1, ie + 1, j, j)
                ptc = put(ptc, is_ - 1, ie + 1, j, j, S(u, is_ - 1, ie + 1, j, j) * S(g.dyc, is_ - 1, ie + 1, j, j) *
                          torch.where(vcv > 0, S(sin4, is_ - 1, ie + 1, j - 1, j - 1), S(sin2, is_ - 1, ie + 1, j, j)))
            vo = Z(u)
            i0, i1 = 2, npx - 1
            vo = put(vo, i0, i1, js - 1, je + 1, (S(v, i0, i1, js - 1, je + 1) - 0.5 * (S(ua, i0 - 1, i1 - 1, js - 1, je + 1) + S(ua, i0, i1, js - 1, je + 1)) * S(g.cosa_u, i0, i1, js - 1, je + 1)) *
                     S(g.dxc, i0, i1, js - 1, je + 1) * S(g.sina_u, i0, i1, js - 1, je + 1))
            for i in (1, npx):
                ucv = S(uc, i, i, js - 1, je + 1)
                vo = put(vo, i, i, js - 1, je + 1, S(v, i, i, js - 1, je + 1) * S(g.dxc, i, i, js - 1, je + 1) *
                         torch.where(ucv > 0, S(sin3, i - 1, i - 1, js - 1, je + 1), S(sin1, i, i, js - 1, je + 1)))
            dpc = (S(vo, C[0], C[1], C[2] - 1, C[3] - 1) - S(vo, *C)) + (S(ptc, C[0] - 1, C[1] - 1, C[2], C[3]) - S(ptc, *C))
            dpc = put(Z(u), *C, dpc).clone()
            dpc[..., 1 + O, 1 + O] = P(dpc, 1, 1) - P(vo, 1, 0)
            dpc[..., 1 + O, npx + O] = P(dpc, npx, 1) - P(vo, npx, 0)
            dpc[..., npy + O, npx + O] = P(dpc, npx, npy) + P(vo, npx, npy)
            dpc[..., npy + O, 1 + O] = P(dpc, 1, npy) + P(vo, 1, npy)
            dpcv = S(g.rarea_c, *C) * S(dpc, *C)
            damp = da_min_c * torch.maximum(d2bg, torch.clamp(dddmp * torch.abs(dpcv * dt), max=0.20))
            ke_damp["0"] = put(Z(u), *C, damp * dpcv)
            divg["0"] = put(Z(u), *C, dpcv)
        if any(n > 0 for n in nord_l):
            nmax = max(nord_l)
            variants = {}
            for nord in sorted(set(n for n in nord_l if n > 0)):
                delpc = put(Z(u), *C, S(divg_d, *C))
                dd = divg_d
                ucw, vcw = None, None
                for n in range(1, nord + 1):
                    nt = nord - n
                    fill_c = nt != 0
                    if fill_c:
                        dd = fill_corners_bgrid(dd, npx, npy, "x")
                    i0, i1, j0, j1 = is_ - 1 - nt, ie + 1 + nt, js - nt, je + 1 + nt
                    vcw = put(Z(u), i0, i1, j0, j1, (S(dd, i0 + 1, i1 + 1, j0, j1) - S(dd, i0, i1, j0, j1)) * S(g.divg_u, i0, i1, j0, j1))
                    if fill_c:
                        dd = fill_corners_bgrid(dd, npx, npy, "y")
                    i0, i1, j0, j1 = is_ - nt, ie + 1 + nt, js - 1 - nt, je + 1 + nt
                    ucw = put(Z(u), i0, i1, j0, j1, (S(dd, i0, i1, j0 + 1, j1 + 1) - S(dd, i0, i1, j0, j1)) * S(g.divg_v, i0, i1, j0, j1))
                    if fill_c:
                        vcw, ucw = fill_corners_dgrid(vcw, ucw, npx, npy, -1.0)
                    i0, i1, j0, j1 = is_ - nt, ie + 1 + nt, js - nt, je + 1 + nt
                    ddn = (S(ucw, i0, i1, j0 - 1, j1 - 1) - S(ucw, i0, i1, j0, j1)) + (S(vcw, i0 - 1, i1 - 1, j0, j1) - S(vcw, i0, i1, j0, j1))
                    ddn = put(Z(u), i0, i1, j0, j1, ddn).clone()
                    ddn[..., 1 + O, 1 + O] = P(ddn, 1, 1) - P(ucw, 1, 0)
                    ddn[..., 1 + O, npx + O] = P(ddn, npx, 1) - P(ucw, npx, 0)
                    ddn[..., npy + O, npx + O] = P(ddn, npx, npy) + P(ucw, npx, npy)
                    ddn[..., npy + O, 1 + O] = P(ddn, 1, npy) + P(ucw, 1, npy)
                    dd = put(Z(u), i0, i1, j0, j1, S(ddn, i0, i1, j0, j1) * S(g.rarea_c, i0, i1, j0, j1))
                if dddmp < 1.e-5:
                    vort = Z(u)
                else:
                    vq = a2b_ord4(wk, g)
                    vort = put(Z(u), *C, abs(dt) * torch.sqrt(S(delpc, *C) ** 2 + S(vq, *C) ** 2))
                dd8 = (da_min_c * d4_bg) ** (nord + 1)
                damp2 = da_min_c * torch.maximum(d2bg, torch.clamp(dddmp * S(vort, *C), max=0.20))
                variants[str(nord)] = put(Z(u), *C, damp2 * S(delpc, *C) + dd8 * S(dd, *C))
                divg[str(nord)] = delpc
            ke_damp.update(variants)
        vort_d = tp.lev_select([str(n) for n in nord_l], ke_damp)
        return vort_d, tp.lev_select([str(n) for n in nord_l], divg)
    if not split_damp:
        vort_d, divg_out = div_damp(prm)
    else:
        (va_, ga_), (vb_, gb_) = div_damp(pp), div_damp(prm)
        vort_d, divg_out = splice(va_, vb_), splice(ga_, gb_)
    ke = put(ke, *C, S(ke, *C) + S(vort_d, *C))

    # ---- vorticity transport and wind update (:1449-1483)
    vort = put(Z(u), isd, ied, jsd, jed, S(wk, isd, ied, jsd, jed) + S(g.f0, isd, ied, jsd, jed))
    fxv, fyv, _ = tp_site(prm["hord_vt"] == pp["hord_vt"], lambda p: tp.fv_tp_2d(vort, crx, cry, p["hord_vt"], xfx, yfx, g, ra_x, ra_y))
    i0, i1, j0, j1 = is_, ie, js, je + 1
    un = S(vt2, i0, i1, j0, j1) + (S(ke, i0, i1, j0, j1) - S(ke, i0 + 1, i1 + 1, j0, j1)) + S(fyv, i0, i1, j0, j1)
    i0, i1, j0, j1 = is_, ie + 1, js, je
    vn = S(ut2, i0, i1, j0, j1) + (S(ke, i0, i1, j0, j1) - S(ke, i0, i1, j0 + 1, j1 + 1)) - S(fxv, i0, i1, j0, j1)
    un_pre, vn_pre = un, vn
    # ---- vorticity damping (:1487-1539)
    dmpv = prm["damp_v"]; nordv = prm["nord_v"]
    if any(d > 1.e-5 for d in dmpv) or any(d > 1.e-5 for d in pp["damp_v"]):
        def del6v(p):       # levels whose damping is off get damp4 = 0: exact zero fluxes there
            d4 = [(p["damp_v"][k] * g.da_min_c) ** (p["nord_v"][k] + 1) if p["damp_v"][k] > 1.e-5 else 0.0 for k in range(K)]
            return del6_by_level([p["nord_v"][k] if p["damp_v"][k] > 1.e-5 else 0 for k in range(K)], d4, wk, g)
        ut3, vt3 = del6v(prm)
        if pp["nord_v"] != nordv or pp["damp_v"] != dmpv:
            # the perturbation always uses its own pair, each side only on the levels where its damping is on (:2436-2451, 2506-2530)
            ua3, va3 = del6v(pp)
            ut3, vt3 = splice(ua3, ut3), splice(va3, vt3)
        un = un + S(vt3, is_, ie, js, je + 1)
        vn = vn - S(ut3, is_, ie + 1, js, je)
    u_new = put(u, is_, ie, js, je + 1, un)
    v_new = put(v, is_, ie + 1, js, je, vn)
    out = dict(delp=delp_new, pt=pt_new, u=u_new, v=v_new, w=w_new, fx=fx, fy=fy, crx=crx, cry=cry, xfx=xfx, yfx=yfx, divg=divg_out)
    # ---- dissipative heating (:938-951 the w part with ke_bg = 0, :1436-1446, :1494-1525).  prm["d_con"]: d_con_k per level
    dcon = prm.get("d_con")
    if dcon is not None and any(d > 1.e-5 for d in dcon):
        assert all(dmpv[k] > 1.e-5 for k in range(K) if dcon[k] > 1.e-5)
        heat0 = torch.zeros_like(delp_new_v)
        if not hydro:
            heat0 = -(dw * (S(w, is_, ie, js, je) + 0.5 * dw))
        i0, i1, j0, j1 = is_, ie, js, je + 1
        ubh = put(Z(u), i0, i1, j0, j1, ((S(vort_d, i0, i1, j0, j1) - S(vort_d, i0 + 1, i1 + 1, j0, j1)) + S(vt3, i0, i1, j0, j1)) * S(g.rdx, i0, i1, j0, j1))
        fyh = put(Z(u), i0, i1, j0, j1, un_pre * S(g.rdx, i0, i1, j0, j1))
        i0, i1, j0, j1 = is_, ie + 1, js, je
        vbh = put(Z(u), i0, i1, j0, j1, ((S(vort_d, i0, i1, j0, j1) - S(vort_d, i0, i1, j0 + 1, j1 + 1)) - S(ut3, i0, i1, j0, j1)) * S(g.rdy, i0, i1, j0, j1))
        fxh = put(Z(u), i0, i1, j0, j1, vn_pre * S(g.rdy, i0, i1, j0, j1))
        C0 = (is_, ie, js, je); CN = (is_, ie, js + 1, je + 1); CE = (is_ + 1, ie + 1, js, je)
        ub0, ub1, fy0, fy1 = S(ubh, *C0), S(ubh, *CN), S(fyh, *C0), S(fyh, *CN)
        vb0, vb1, fx0, fx1 = S(vbh, *C0), S(vbh, *CE), S(fxh, *C0), S(fxh, *CE)
        u2 = fy0 + fy1; du2 = ub0 + ub1; v2 = fx0 + fx1; dv2 = vb0 + vb1
        dk = _lv(dcon)
        hs = delp_new_v * (heat0 - 0.25 * dk * S(g.rsin2, *C0) * ((ub0 ** 2 + ub1 ** 2 + vb0 ** 2 + vb1 ** 2) +
                                                               2. * (fy0 * ub0 + fy1 * ub1 + fx0 * vb0 + fx1 * vb1) -
                                                               S(g.cosa_s, *C0) * (u2 * dv2 + v2 * du2 + du2 * dv2)))
        out["heat"] = put(Z(u), *C0, torch.where(dk > 1.e-5, hs, heat0))
    return out
