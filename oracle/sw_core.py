"""ORACLE (test infrastructure only).  sw_core: C-grid half step c_sw and D-grid step d_sw
restated in torch float64 (differentiable: TL = jvp, AD = vjp) from

  model/sw_core_nlm.F90   c_sw :77-486, d_sw :492-1545, del6_vt_flux :1547-1659,
                          divergence_corner :1661-1766, xtp_u :1970, ytp_v :2312,
                          d2a2c_vect :2746-3085, edge_interpolate4 :3087, fill_4corners :3235
  model_tlmadm/sw_core_tlm.F90  (Tapenade primal copies; iord = 1 branch of XTP_U/YTP_V :7320)

Whole-tile decomposition (is=js=1, ie=je=N): every tile has all four cube corners and
all four edges.  grid_type = 0, non-nested, non-stretched.  Linear orders only (1, 2).

parity: pinned by tests/test_ref_tlm.py (C_SW_TLM, D2A2C_VECT_TLM, DIVERGENCE_CORNER_TLM transliterated: bit for bit / 6e-15) and by the
reference's DYN_CORE_TLM executed (tests/test_ref_golden.py).  Arrays [6, K, NY, NX]; (i,j) at [..., j+2, i+2].
"""
import torch
from .cubed_sphere import R, NG, fill_4corners, copy_corners, fill_corners_bgrid, fill_corners_dgrid
from . import tp_core as tp
from .a2b_edge import a2b_ord4

A1, A2 = 0.5625, -0.0625                    # 4-pt Lagrange interpolation (sw_core_nlm.F90:46-47)
C1, C2, C3 = -2. / 14., 11. / 14., 5. / 14.   # volume-conserving cubic, 2nd-deriv=0 at end point
P1, P2 = 7. / 12., -1. / 12.
BIG = 1.e30   # stands for big_number in "utmp(:,:) = big_number": never read on valid paths; use 0
O = NG - 1


def Z(a):
    return torch.zeros_like(a)


def put(a, i0, i1, j0, j1, v):
    a = a.clone()
    a[..., R(j0, j1), R(i0, i1)] = v
    return a


def S(a, i0, i1, j0, j1):
    return a[..., R(j0, j1), R(i0, i1)]


def P(a, i, j):
    """single point (keeps [6,K])"""
    return a[..., j + O, i + O]


def sg(g, k):
    return g.sin_sg[..., k], g.cos_sg[..., k]


def edge_interpolate4(u1, u2, u3, u4, d1, d2, d3, d4):
    t1 = d1 + d2
    t2 = d3 + d4
    return 0.5 * (((t1 + d2) * u2 - d2 * u1) / t1 + ((t2 + d3) * u3 - d3 * u4) / t2)


# ------------------------------------------------------------------------------------
def d2a2c_vect(u, v, g, dord4=True):
    """sw_core_nlm.F90:2746-3085.  returns ua, va, uc, vc, ut, vt (full arrays)"""
    N, npx, npy = g.N, g.npx, g.npy
    is_, ie, js, je = 1, N, 1, N
    isd, ied, jsd, jed = g.isd, g.ied, g.jsd, g.jed
    npt = 4
    idd = 1 if dord4 else 0
    sin1, _ = sg(g, 1); sin2, _ = sg(g, 2); sin3, _ = sg(g, 3); sin4, _ = sg(g, 4)
    utmp = Z(u); vtmp = Z(u)
    # interior 4th-order D->A
    j0, j1 = max(npt, js - 1), min(npy - npt, je + 1)
    i0, i1 = max(npt, isd), min(npx - npt, ied)
    utmp = put(utmp, i0, i1, j0, j1, A2 * (S(u, i0, i1, j0 - 1, j1 - 1) + S(u, i0, i1, j0 + 2, j1 + 2)) +
               A1 * (S(u, i0, i1, j0, j1) + S(u, i0, i1, j0 + 1, j1 + 1)))
    j0, j1 = max(npt, jsd), min(npy - npt, jed)
    i0, i1 = max(npt, is_ - 1), min(npx - npt, ie + 1)
    vtmp = put(vtmp, i0, i1, j0, j1, A2 * (S(v, i0 - 1, i1 - 1, j0, j1) + S(v, i0 + 2, i1 + 2, j0, j1)) +
               A1 * (S(v, i0, i1, j0, j1) + S(v, i0 + 1, i1 + 1, j0, j1)))

    def lin(i0, i1, j0, j1):
        nonlocal utmp, vtmp
        utmp = put(utmp, i0, i1, j0, j1, 0.5 * (S(u, i0, i1, j0, j1) + S(u, i0, i1, j0 + 1, j1 + 1)))
        vtmp = put(vtmp, i0, i1, j0, j1, 0.5 * (S(v, i0, i1, j0, j1) + S(v, i0 + 1, i1 + 1, j0, j1)))
    lin(isd, ied, jsd, npt - 1)
    lin(isd, ied, npy - npt + 1, jed)
    lin(isd, npt - 1, max(npt, jsd), min(npy - npt, jed))
    lin(npx - npt + 1, ied, max(npt, jsd), min(npy - npt, jed))

    ua = Z(u); va = Z(u)
    i0, i1, j0, j1 = is_ - 1 - idd, ie + 1 + idd, js - 1 - idd, je + 1 + idd
    cs = S(g.cosa_s, i0, i1, j0, j1); r2 = S(g.rsin2, i0, i1, j0, j1)
    ua = put(ua, i0, i1, j0, j1, (S(utmp, i0, i1, j0, j1) - S(vtmp, i0, i1, j0, j1) * cs) * r2)
    va = put(va, i0, i1, j0, j1, (S(vtmp, i0, i1, j0, j1) - S(utmp, i0, i1, j0, j1) * cs) * r2)

    # A -> C, x direction
    ut0 = utmp.clone()
    for i in (-2, -1, 0):
        ut0[..., 0 + O, i + O] = -P(vtmp, 0, 1 - i)                   # sw
        ut0[..., npy + O, i + O] = P(vtmp, 0, je + i)                 # nw
    for i in (0, 1, 2):
        ut0[..., 0 + O, npx + i + O] = P(vtmp, npx, i + 1)            # se
        ut0[..., npy + O, npx + i + O] = -P(vtmp, npx, je - i)        # ne
    utmp = ut0
    uc = Z(u); ut = Z(u)
    ifirst, ilast = max(3, is_ - 1), min(npx - 2, ie + 2)
    j0, j1 = js - 1, je + 1
    ucv = A2 * (S(utmp, ifirst - 2, ilast - 2, j0, j1) + S(utmp, ifirst + 1, ilast + 1, j0, j1)) + \
        A1 * (S(utmp, ifirst - 1, ilast - 1, j0, j1) + S(utmp, ifirst, ilast, j0, j1))
    uc = put(uc, ifirst, ilast, j0, j1, ucv)
    ut = put(ut, ifirst, ilast, j0, j1, (ucv - S(v, ifirst, ilast, j0, j1) * S(g.cosa_u, ifirst, ilast, j0, j1)) *
             S(g.rsin_u, ifirst, ilast, j0, j1))
    ua = ua.clone()
    ua[..., 0 + O, -1 + O] = -P(va, 0, 2); ua[..., 0 + O, 0 + O] = -P(va, 0, 1)                     # sw
    ua[..., 0 + O, npx + O] = P(va, npx, 1); ua[..., 0 + O, npx + 1 + O] = P(va, npx, 2)            # se
    ua[..., npy + O, npx + O] = -P(va, npx, npy - 1); ua[..., npy + O, npx + 1 + O] = -P(va, npx, npy - 2)  # ne
    ua[..., npy + O, -1 + O] = P(va, 0, npy - 2); ua[..., npy + O, 0 + O] = P(va, 0, npy - 1)       # nw

    def col(a, i):
        return a[..., R(j0, j1), i + O]
    uc = uc.clone(); ut = ut.clone()
    # west edge
    uc[..., R(j0, j1), 0 + O] = C1 * col(utmp, -2) + C2 * col(utmp, -1) + C3 * col(utmp, 0)
    ut1 = edge_interpolate4(col(ua, -1), col(ua, 0), col(ua, 1), col(ua, 2),
                            col(g.dxa, -1), col(g.dxa, 0), col(g.dxa, 1), col(g.dxa, 2))
    ut[..., R(j0, j1), 1 + O] = ut1
    uc[..., R(j0, j1), 1 + O] = torch.where(ut1 > 0., ut1 * col(sin3, 0), ut1 * col(sin1, 1))
    uc[..., R(j0, j1), 2 + O] = C1 * col(utmp, 3) + C2 * col(utmp, 2) + C3 * col(utmp, 1)
    ut[..., R(j0, j1), 0 + O] = (col(uc, 0) - col(v, 0) * col(g.cosa_u, 0)) * col(g.rsin_u, 0)
    ut[..., R(j0, j1), 2 + O] = (col(uc, 2) - col(v, 2) * col(g.cosa_u, 2)) * col(g.rsin_u, 2)
    # east edge
    uc[..., R(j0, j1), npx - 1 + O] = C1 * col(utmp, npx - 3) + C2 * col(utmp, npx - 2) + C3 * col(utmp, npx - 1)
    utn = edge_interpolate4(col(ua, npx - 2), col(ua, npx - 1), col(ua, npx), col(ua, npx + 1),
                            col(g.dxa, npx - 2), col(g.dxa, npx - 1), col(g.dxa, npx), col(g.dxa, npx + 1))
    ut[..., R(j0, j1), npx + O] = utn
    uc[..., R(j0, j1), npx + O] = torch.where(utn > 0., utn * col(sin3, npx - 1), utn * col(sin1, npx))
    uc[..., R(j0, j1), npx + 1 + O] = C3 * col(utmp, npx) + C2 * col(utmp, npx + 1) + C1 * col(utmp, npx + 2)
    ut[..., R(j0, j1), npx - 1 + O] = (col(uc, npx - 1) - col(v, npx - 1) * col(g.cosa_u, npx - 1)) * col(g.rsin_u, npx - 1)
    ut[..., R(j0, j1), npx + 1 + O] = (col(uc, npx + 1) - col(v, npx + 1) * col(g.cosa_u, npx + 1)) * col(g.rsin_u, npx + 1)

    # A -> C, y direction
    vt0 = vtmp.clone()
    for j in (-2, -1, 0):
        vt0[..., j + O, 0 + O] = -P(utmp, 1 - j, 0)                   # sw
        vt0[..., j + O, npx + O] = P(utmp, ie + j, 0)                 # se
    for j in (0, 1, 2):
        vt0[..., npy + j + O, 0 + O] = P(utmp, j + 1, npy)            # nw
        vt0[..., npy + j + O, npx + O] = -P(utmp, ie - j, npy)        # ne
    vtmp = vt0
    va = va.clone()
    va[..., -1 + O, 0 + O] = -P(ua, 2, 0); va[..., 0 + O, 0 + O] = -P(ua, 1, 0)                     # sw
    va[..., 0 + O, npx + O] = P(ua, npx - 1, 0); va[..., -1 + O, npx + O] = P(ua, npx - 2, 0)       # se
    va[..., npy + O, npx + O] = -P(ua, npx - 1, npy); va[..., npy + 1 + O, npx + O] = -P(ua, npx - 2, npy)  # ne
    va[..., npy + O, 0 + O] = P(ua, 1, npy); va[..., npy + 1 + O, 0 + O] = P(ua, 2, npy)            # nw

    vc = Z(u); vt = Z(u)
    i0, i1 = is_ - 1, ie + 1

    def row(a, j):
        return a[..., j + O, R(i0, i1)]
    for j in range(js - 1, je + 3):
        if j == 1 or j == npy:
            vtj = edge_interpolate4(row(va, j - 2), row(va, j - 1), row(va, j), row(va, j + 1),
                                    row(g.dya, j - 2), row(g.dya, j - 1), row(g.dya, j), row(g.dya, j + 1))
            vt[..., j + O, R(i0, i1)] = vtj
            vc[..., j + O, R(i0, i1)] = torch.where(vtj > 0., vtj * row(sin4, j - 1), vtj * row(sin2, j))
        else:
            if j == 0 or j == npy - 1:
                vcj = C1 * row(vtmp, j - 2) + C2 * row(vtmp, j - 1) + C3 * row(vtmp, j)
            elif j == 2 or j == npy + 1:
                vcj = C1 * row(vtmp, j + 1) + C2 * row(vtmp, j) + C3 * row(vtmp, j - 1)
            else:
                vcj = A2 * (row(vtmp, j - 2) + row(vtmp, j + 1)) + A1 * (row(vtmp, j - 1) + row(vtmp, j))
            vc[..., j + O, R(i0, i1)] = vcj
            vt[..., j + O, R(i0, i1)] = (vcj - row(u, j) * row(g.cosa_v, j)) * row(g.rsin_v, j)
    return ua, va, uc, vc, ut, vt


def divergence_corner(u, v, ua, va, g):
    """sw_core_nlm.F90:1661-1766"""
    N, npx, npy = g.N, g.npx, g.npy
    is_, ie, js, je = 1, N, 1, N
    sin1, cos1 = sg(g, 1); sin2, cos2 = sg(g, 2); sin3, cos3 = sg(g, 3); sin4, cos4 = sg(g, 4)
    uf = Z(u); vf = Z(u)
    i0, i1, j0, j1 = is_ - 1, ie + 1, js, je + 1
    ufv = (S(u, i0, i1, j0, j1) - 0.25 * (S(va, i0, i1, j0 - 1, j1 - 1) + S(va, i0, i1, j0, j1)) *
           (S(cos4, i0, i1, j0 - 1, j1 - 1) + S(cos2, i0, i1, j0, j1))) * S(g.dyc, i0, i1, j0, j1) * 0.5 * \
        (S(sin4, i0, i1, j0 - 1, j1 - 1) + S(sin2, i0, i1, j0, j1))
    uf = put(uf, i0, i1, j0, j1, ufv)
    for j in (1, npy):
        uf = put(uf, i0, i1, j, j, S(u, i0, i1, j, j) * S(g.dyc, i0, i1, j, j) * 0.5 *
                 (S(sin4, i0, i1, j - 1, j - 1) + S(sin2, i0, i1, j, j)))
    i0, i1, j0, j1 = 2, npx - 1, js - 1, je + 1
    vfv = (S(v, i0, i1, j0, j1) - 0.25 * (S(ua, i0 - 1, i1 - 1, j0, j1) + S(ua, i0, i1, j0, j1)) *
           (S(cos3, i0 - 1, i1 - 1, j0, j1) + S(cos1, i0, i1, j0, j1))) * S(g.dxc, i0, i1, j0, j1) * 0.5 * \
        (S(sin3, i0 - 1, i1 - 1, j0, j1) + S(sin1, i0, i1, j0, j1))
    vf = put(vf, i0, i1, j0, j1, vfv)
    for i in (1, npx):
        vf = put(vf, i, i, j0, j1, S(v, i, i, j0, j1) * S(g.dxc, i, i, j0, j1) * 0.5 *
                 (S(sin3, i - 1, i - 1, j0, j1) + S(sin1, i, i, j0, j1)))
    i0, i1, j0, j1 = is_, ie + 1, js, je + 1
    d = (S(vf, i0, i1, j0 - 1, j1 - 1) - S(vf, i0, i1, j0, j1)) + (S(uf, i0 - 1, i1 - 1, j0, j1) - S(uf, i0, i1, j0, j1))
    dv = put(Z(u), i0, i1, j0, j1, d)
    dv = dv.clone()
    dv[..., 1 + O, 1 + O] = P(dv, 1, 1) - P(vf, 1, 0)
    dv[..., 1 + O, npx + O] = P(dv, npx, 1) - P(vf, npx, 0)
    dv[..., npy + O, npx + O] = P(dv, npx, npy) + P(vf, npx, npy)
    dv[..., npy + O, 1 + O] = P(dv, 1, npy) + P(vf, 1, npy)
    return put(Z(u), i0, i1, j0, j1, S(g.rarea_c, i0, i1, j0, j1) * S(dv, i0, i1, j0, j1))


def c_sw(delp, pt, u, v, w, g, dt2, hydrostatic, nord=1):
    """sw_core_nlm.F90:77-486.  Returns dict(delpc, ptc, wc, uc, vc, ua, va, ut, vt, divg_d)
    (ut, vt are the Courant-number-like transporting winds after :157-174)."""
    N, npx, npy = g.N, g.npx, g.npy
    is_, ie, js, je = 1, N, 1, N
    sin1, cos1 = sg(g, 1); sin2, cos2 = sg(g, 2); sin3, cos3 = sg(g, 3); sin4, cos4 = sg(g, 4)
    ua, va, uc, vc, ut, vt = d2a2c_vect(u, v, g, True)
    divg_d = divergence_corner(u, v, ua, va, g) if nord > 0 else Z(u)

    i0, i1, j0, j1 = is_ - 1, ie + 2, js - 1, je + 1
    utv = S(ut, i0, i1, j0, j1)
    utn = dt2 * utv * S(g.dy, i0, i1, j0, j1) * torch.where(utv > 0., S(sin3, i0 - 1, i1 - 1, j0, j1), S(sin1, i0, i1, j0, j1))
    ut = put(ut, i0, i1, j0, j1, utn)
    i0, i1, j0, j1 = is_ - 1, ie + 1, js - 1, je + 2
    vtv = S(vt, i0, i1, j0, j1)
    vtn = dt2 * vtv * S(g.dx, i0, i1, j0, j1) * torch.where(vtv > 0., S(sin4, i0, i1, j0 - 1, j1 - 1), S(sin2, i0, i1, j0, j1))
    vt = put(vt, i0, i1, j0, j1, vtn)

    # ---- x fluxes (first-order upwind)
    delp1 = fill_4corners(delp, npx, npy, 1); pt1 = fill_4corners(pt, npx, npy, 1)
    w1 = None if hydrostatic else fill_4corners(w, npx, npy, 1)
    i0, i1, j0, j1 = is_ - 1, ie + 2, js - 1, je + 1
    utv = S(ut, i0, i1, j0, j1)
    up = utv > 0.
    fx1 = utv * torch.where(up, S(delp1, i0 - 1, i1 - 1, j0, j1), S(delp1, i0, i1, j0, j1))
    fx = fx1 * torch.where(up, S(pt1, i0 - 1, i1 - 1, j0, j1), S(pt1, i0, i1, j0, j1))
    if not hydrostatic:
        fx2 = fx1 * torch.where(up, S(w1, i0 - 1, i1 - 1, j0, j1), S(w1, i0, i1, j0, j1))
    # ---- y fluxes
    delp2 = fill_4corners(delp1, npx, npy, 2); pt2 = fill_4corners(pt1, npx, npy, 2)
    w2 = None if hydrostatic else fill_4corners(w1, npx, npy, 2)
    i0, i1, j0, j1 = is_ - 1, ie + 1, js - 1, je + 2
    vtv = S(vt, i0, i1, j0, j1)
    up = vtv > 0.
    fy1 = vtv * torch.where(up, S(delp2, i0, i1, j0 - 1, j1 - 1), S(delp2, i0, i1, j0, j1))
    fy = fy1 * torch.where(up, S(pt2, i0, i1, j0 - 1, j1 - 1), S(pt2, i0, i1, j0, j1))
    if not hydrostatic:
        fy2 = fy1 * torch.where(up, S(w2, i0, i1, j0 - 1, j1 - 1), S(w2, i0, i1, j0, j1))
    # fx* shape [.., nj=je+1-(js-1)+1, ni=ie+2-(is-1)+1]; fy* one more row, one fewer column
    i0, i1, j0, j1 = is_ - 1, ie + 1, js - 1, je + 1
    ra = S(g.rarea, i0, i1, j0, j1)
    dp_, pt_ = S(delp2, i0, i1, j0, j1), S(pt2, i0, i1, j0, j1)
    delpc_v = dp_ + ((fx1[..., :, :-1] - fx1[..., :, 1:]) + (fy1[..., :-1, :] - fy1[..., 1:, :])) * ra
    ptc_v = (pt_ * dp_ + ((fx[..., :, :-1] - fx[..., :, 1:]) + (fy[..., :-1, :] - fy[..., 1:, :])) * ra) / delpc_v
    delpc = put(Z(u), i0, i1, j0, j1, delpc_v)
    ptc = put(Z(u), i0, i1, j0, j1, ptc_v)
    wc = Z(u)
    if not hydrostatic:
        wc_v = (S(w2, i0, i1, j0, j1) * dp_ + ((fx2[..., :, :-1] - fx2[..., :, 1:]) + (fy2[..., :-1, :] - fy2[..., 1:, :])) * ra) / delpc_v
        wc = put(wc, i0, i1, j0, j1, wc_v)

    # ---- kinetic energy (:314-364)
    uav = S(ua, i0, i1, j0, j1); vav = S(va, i0, i1, j0, j1)
    ke_pos = S(uc, i0, i1, j0, j1).clone()
    ke_neg = S(uc, i0 + 1, i1 + 1, j0, j1).clone()
    # i == 1 / i == npx columns of the "ua>0" branch; i == 0 / npx-1 of the other
    def cidx(i): return i - i0
    ke_pos[..., :, cidx(1)] = S(uc, 1, 1, j0, j1)[..., 0] * S(sin1, 1, 1, j0, j1)[..., 0] + S(v, 1, 1, j0, j1)[..., 0] * S(cos1, 1, 1, j0, j1)[..., 0]
    ke_pos[..., :, cidx(npx)] = S(uc, npx, npx, j0, j1)[..., 0] * S(sin1, npx, npx, j0, j1)[..., 0] + S(v, npx, npx, j0, j1)[..., 0] * S(cos1, npx, npx, j0, j1)[..., 0]
    ke_neg[..., :, cidx(0)] = S(uc, 1, 1, j0, j1)[..., 0] * S(sin3, 0, 0, j0, j1)[..., 0] + S(v, 1, 1, j0, j1)[..., 0] * S(cos3, 0, 0, j0, j1)[..., 0]
    ke_neg[..., :, cidx(npx - 1)] = S(uc, npx, npx, j0, j1)[..., 0] * S(sin3, npx - 1, npx - 1, j0, j1)[..., 0] + S(v, npx, npx, j0, j1)[..., 0] * S(cos3, npx - 1, npx - 1, j0, j1)[..., 0]
    ke = torch.where(uav > 0., ke_pos, ke_neg)
    vo_pos = S(vc, i0, i1, j0, j1).clone()
    vo_neg = S(vc, i0, i1, j0 + 1, j1 + 1).clone()
    def ridx(j): return j - j0
    vo_pos[..., ridx(1), :] = S(vc, i0, i1, 1, 1)[..., 0, :] * S(sin2, i0, i1, 1, 1)[..., 0, :] + S(u, i0, i1, 1, 1)[..., 0, :] * S(cos2, i0, i1, 1, 1)[..., 0, :]
    vo_pos[..., ridx(npy), :] = S(vc, i0, i1, npy, npy)[..., 0, :] * S(sin2, i0, i1, npy, npy)[..., 0, :] + S(u, i0, i1, npy, npy)[..., 0, :] * S(cos2, i0, i1, npy, npy)[..., 0, :]
    vo_neg[..., ridx(0), :] = S(vc, i0, i1, 1, 1)[..., 0, :] * S(sin4, i0, i1, 0, 0)[..., 0, :] + S(u, i0, i1, 1, 1)[..., 0, :] * S(cos4, i0, i1, 0, 0)[..., 0, :]
    vo_neg[..., ridx(npy - 1), :] = S(vc, i0, i1, npy, npy)[..., 0, :] * S(sin4, i0, i1, npy - 1, npy - 1)[..., 0, :] + S(u, i0, i1, npy, npy)[..., 0, :] * S(cos4, i0, i1, npy - 1, npy - 1)[..., 0, :]
    vort = torch.where(vav > 0., vo_pos, vo_neg)
    dt4 = 0.5 * dt2
    ke = put(Z(u), i0, i1, j0, j1, dt4 * (uav * ke + vav * vort))

    # ---- circulation and absolute vorticity on corners (:370-401)
    fxc = uc * g.dxc      # (is:ie+1, js-1:je+1)
    fyc = vc * g.dyc      # (is-1:ie+1, js:je+1)
    i0, i1, j0, j1 = is_, ie + 1, js, je + 1
    vo = (S(fxc, i0, i1, j0 - 1, j1 - 1) - S(fxc, i0, i1, j0, j1)) + (S(fyc, i0, i1, j0, j1) - S(fyc, i0 - 1, i1 - 1, j0, j1))
    vo = put(Z(u), i0, i1, j0, j1, vo).clone()
    vo[..., 1 + O, 1 + O] = P(vo, 1, 1) + P(fyc, 0, 1)
    vo[..., 1 + O, npx + O] = P(vo, npx, 1) - P(fyc, npx, 1)
    vo[..., npy + O, npx + O] = P(vo, npx, npy) - P(fyc, npx, npy)
    vo[..., npy + O, 1 + O] = P(vo, 1, npy) + P(fyc, 0, npy)
    vo = put(Z(u), i0, i1, j0, j1, S(g.fC, i0, i1, j0, j1) + S(g.rarea_c, i0, i1, j0, j1) * S(vo, i0, i1, j0, j1))

    # ---- vorticity transport (:434-472) and wind update (:475-484)
    i0, i1, j0, j1 = is_, ie + 1, js, je
    fy1v = dt2 * (S(v, i0, i1, j0, j1) - S(uc, i0, i1, j0, j1) * S(g.cosa_u, i0, i1, j0, j1)) / S(g.sina_u, i0, i1, j0, j1)
    fy1v = fy1v.clone()
    fy1v[..., :, 0] = dt2 * S(v, 1, 1, j0, j1)[..., 0]
    fy1v[..., :, -1] = dt2 * S(v, npx, npx, j0, j1)[..., 0]
    fyv = torch.where(fy1v > 0., S(vo, i0, i1, j0, j1), S(vo, i0, i1, j0 + 1, j1 + 1))
    ucn = S(uc, i0, i1, j0, j1) + fy1v * fyv + S(g.rdxc, i0, i1, j0, j1) * (S(ke, i0 - 1, i1 - 1, j0, j1) - S(ke, i0, i1, j0, j1))
    i0, i1, j0, j1 = is_, ie, js, je + 1
    fx1v = dt2 * (S(u, i0, i1, j0, j1) - S(vc, i0, i1, j0, j1) * S(g.cosa_v, i0, i1, j0, j1)) / S(g.sina_v, i0, i1, j0, j1)
    fx1v = fx1v.clone()
    fx1v[..., 0, :] = dt2 * S(u, i0, i1, 1, 1)[..., 0, :]
    fx1v[..., -1, :] = dt2 * S(u, i0, i1, npy, npy)[..., 0, :]
    fxv = torch.where(fx1v > 0., S(vo, i0, i1, j0, j1), S(vo, i0 + 1, i1 + 1, j0, j1))
    vcn = S(vc, i0, i1, j0, j1) - fx1v * fxv + S(g.rdyc, i0, i1, j0, j1) * (S(ke, i0, i1, j0 - 1, j1 - 1) - S(ke, i0, i1, j0, j1))
    uc = put(uc, is_, ie + 1, js, je, ucn)
    vc = put(vc, is_, ie, js, je + 1, vcn)
    return dict(delpc=delpc, ptc=ptc, wc=wc, uc=uc, vc=vc, ua=ua, va=va, ut=ut, vt=vt, divg_d=divg_d)
