"""ORACLE (test infrastructure only).  a2b_ord4: 4th-order A-grid -> B-grid (cell corner)
interpolation, restated from model/a2b_edge_nlm.F90:49-329 (extrap_corner :800-810;
TL model_tlmadm/a2b_edge_tlm.F90:546, AD a2b_edge_adm.F90:72).  Linear in qin: the TL is
the same operator and the AD its transpose (obtained here by autograd).

parity: pinned by tests/test_ref_tlm.py (A2B_ORD4_TLM + EXTRAP_CORNER_TLM transliterated, 1e-16) and through D_SW_TLM / NH_P_GRAD_TLM.  Whole tile: is=js=1, ie=je=N.
"""
import numpy as np
import torch
from .cubed_sphere import R, NG
from . import grid as G

A1, A2 = 0.5625, -0.0625
B1, B2 = 7. / 12., -1. / 12.
C1, C2 = 2. / 3., -1. / 6.
R3 = 1. / 3.
O = NG - 1


def _corner_coef(g):
    """x1/(x2-x1) for the 3 extrapolations at each of the 4 cube corners: dict corner -> [(p1 idx, p2 idx, coef[6])]"""
    if hasattr(g, "_a2b_corner"):
        return g._a2b_corner
    npx, npy = g.npx, g.npy
    grid = g.grid[:, 0]; agrid = g.agrid[:, 0]
    def GP(i, j): return grid[:, j + O, i + O, :]
    def AP(i, j): return agrid[:, j + O, i + O, :]
    def gcd(q1, q2):   # great_circle_dist (model/fv_grid_utils_nlm.F90:1967), torch version
        return 2.0 * torch.asin(torch.sqrt(torch.sin((q1[..., 1] - q2[..., 1]) / 2.0) ** 2 +
                                           torch.cos(q1[..., 1]) * torch.cos(q2[..., 1]) * torch.sin((q1[..., 0] - q2[..., 0]) / 2.0) ** 2))
    spec = {
        (1, 1): [((1, 1), (2, 2)), ((0, 1), (-1, 2)), ((1, 0), (2, -1))],
        (npx, 1): [((npx - 1, 1), (npx - 2, 2)), ((npx - 1, 0), (npx - 2, -1)), ((npx, 1), (npx + 1, 2))],
        (npx, npy): [((npx - 1, npy - 1), (npx - 2, npy - 2)), ((npx, npy - 1), (npx + 1, npy - 2)), ((npx - 1, npy), (npx - 2, npy + 1))],
        (1, npy): [((1, npy - 1), (2, npy - 2)), ((0, npy - 1), (-1, npy - 2)), ((1, npy), (2, npy + 1))],
    }
    out = {}
    for c, lst in spec.items():
        p0 = GP(*c)
        ent = []
        for (a, b) in lst:
            x1 = gcd(AP(*a), p0); x2 = gcd(AP(*b), p0)
            ent.append((a, b, (x1 / (x2 - x1)).reshape(6, 1)))
        out[c] = ent
    g._a2b_corner = out
    return out


def a2b_ord4(qin, g):
    """returns qout (full array, valid on 1..npx, 1..npy)"""
    N, npx, npy = g.N, g.npx, g.npy
    is_, ie, js, je = 1, N, 1, N
    dxa, dya = g.dxa, g.dya

    def Pq(a, i, j): return a[..., j + O, i + O]
    def S(a, i0, i1, j0, j1): return a[..., R(j0, j1), R(i0, i1)]
    def put(a, i0, i1, j0, j1, v):
        a = a.clone(); a[..., R(j0, j1), R(i0, i1)] = v; return a

    qout = torch.zeros_like(qin)
    for c, ent in _corner_coef(g).items():
        acc = 0.
        for (a, b, coef) in ent:
            q1 = Pq(qin, *a); q2 = Pq(qin, *b)
            acc = acc + (q1 + coef * (q1 - q2))
        qout = qout.clone()
        qout[..., c[1] + O, c[0] + O] = acc * R3

    # ---- qx
    qx = torch.zeros_like(qin)
    j0, j1 = 1, npy - 1
    i0, i1 = 3, npx - 2
    qx = put(qx, i0, i1, j0, j1, B2 * (S(qin, i0 - 2, i1 - 2, j0, j1) + S(qin, i0 + 1, i1 + 1, j0, j1)) +
             B1 * (S(qin, i0 - 1, i1 - 1, j0, j1) + S(qin, i0, i1, j0, j1)))
    def colq(a, i, ja=1, jb=None): return a[..., R(ja, npy - 1 if jb is None else jb), i + O]
    # west
    q2w = (colq(qin, 0) * colq(dxa, 1) + colq(qin, 1) * colq(dxa, 0)) / (colq(dxa, 0) + colq(dxa, 1))   # j = 1..npy-1
    ew = g.edge_w[..., R(2, npy - 1)]
    qout = qout.clone()
    qout[..., R(2, npy - 1), 1 + O] = ew * q2w[..., :-1] + (1. - ew) * q2w[..., 1:]
    g_in = colq(dxa, 2) / colq(dxa, 1); g_ou = colq(dxa, -1) / colq(dxa, 0)
    qx1 = 0.5 * (((2. + g_in) * colq(qin, 1) - colq(qin, 2)) / (1. + g_in) + ((2. + g_ou) * colq(qin, 0) - colq(qin, -1)) / (1. + g_ou))
    qx2 = (3. * (g_in * colq(qin, 1) + colq(qin, 2)) - (g_in * qx1 + colq(qx, 3))) / (2. + 2. * g_in)
    qx = qx.clone(); qx[..., R(1, npy - 1), 1 + O] = qx1; qx[..., R(1, npy - 1), 2 + O] = qx2
    # east
    q2e = (colq(qin, npx - 1) * colq(dxa, npx) + colq(qin, npx) * colq(dxa, npx - 1)) / (colq(dxa, npx - 1) + colq(dxa, npx))
    ee = g.edge_e[..., R(2, npy - 1)]
    qout[..., R(2, npy - 1), npx + O] = ee * q2e[..., :-1] + (1. - ee) * q2e[..., 1:]
    g_in = colq(dxa, npx - 2) / colq(dxa, npx - 1); g_ou = colq(dxa, npx + 1) / colq(dxa, npx)
    qxn = 0.5 * (((2. + g_in) * colq(qin, npx - 1) - colq(qin, npx - 2)) / (1. + g_in) + ((2. + g_ou) * colq(qin, npx) - colq(qin, npx + 1)) / (1. + g_ou))
    qxm = (3. * (colq(qin, npx - 2) + g_in * colq(qin, npx - 1)) - (g_in * qxn + colq(qx, npx - 2))) / (2. + 2. * g_in)
    qx = qx.clone(); qx[..., R(1, npy - 1), npx + O] = qxn; qx[..., R(1, npy - 1), npx - 1 + O] = qxm

    # ---- qy
    qy = torch.zeros_like(qin)
    j0, j1 = 3, npy - 2
    i0, i1 = 1, npx - 1
    qy = put(qy, i0, i1, j0, j1, B2 * (S(qin, i0, i1, j0 - 2, j1 - 2) + S(qin, i0, i1, j0 + 1, j1 + 1)) +
             B1 * (S(qin, i0, i1, j0 - 1, j1 - 1) + S(qin, i0, i1, j0, j1)))
    def rowq(a, j): return a[..., j + O, R(1, npx - 1)]
    q1s = (rowq(qin, 0) * rowq(dya, 1) + rowq(qin, 1) * rowq(dya, 0)) / (rowq(dya, 0) + rowq(dya, 1))
    es = g.edge_s[..., R(2, npx - 1)]
    qout[..., 1 + O, R(2, npx - 1)] = es * q1s[..., :-1] + (1. - es) * q1s[..., 1:]
    g_in = rowq(dya, 2) / rowq(dya, 1); g_ou = rowq(dya, -1) / rowq(dya, 0)
    qy1 = 0.5 * (((2. + g_in) * rowq(qin, 1) - rowq(qin, 2)) / (1. + g_in) + ((2. + g_ou) * rowq(qin, 0) - rowq(qin, -1)) / (1. + g_ou))
    qy2 = (3. * (g_in * rowq(qin, 1) + rowq(qin, 2)) - (g_in * qy1 + rowq(qy, 3))) / (2. + 2. * g_in)
    qy = qy.clone(); qy[..., 1 + O, R(1, npx - 1)] = qy1; qy[..., 2 + O, R(1, npx - 1)] = qy2
    q1n = (rowq(qin, npy - 1) * rowq(dya, npy) + rowq(qin, npy) * rowq(dya, npy - 1)) / (rowq(dya, npy - 1) + rowq(dya, npy))
    en = g.edge_n[..., R(2, npx - 1)]
    qout[..., npy + O, R(2, npx - 1)] = en * q1n[..., :-1] + (1. - en) * q1n[..., 1:]
    g_in = rowq(dya, npy - 2) / rowq(dya, npy - 1); g_ou = rowq(dya, npy + 1) / rowq(dya, npy)
    qyn = 0.5 * (((2. + g_in) * rowq(qin, npy - 1) - rowq(qin, npy - 2)) / (1. + g_in) + ((2. + g_ou) * rowq(qin, npy) - rowq(qin, npy + 1)) / (1. + g_ou))
    qym = (3. * (rowq(qin, npy - 2) + g_in * rowq(qin, npy - 1)) - (g_in * qyn + rowq(qy, npy - 2))) / (2. + 2. * g_in)
    qy = qy.clone(); qy[..., npy + O, R(1, npx - 1)] = qyn; qy[..., npy - 1 + O, R(1, npx - 1)] = qym

    # ---- qxx, qyy, average
    qxx = torch.zeros_like(qin)
    j0, j1, i0, i1 = 3, npy - 2, 2, npx - 1
    qxx = put(qxx, i0, i1, j0, j1, A2 * (S(qx, i0, i1, j0 - 2, j1 - 2) + S(qx, i0, i1, j0 + 1, j1 + 1)) +
              A1 * (S(qx, i0, i1, j0 - 1, j1 - 1) + S(qx, i0, i1, j0, j1)))
    qxx = put(qxx, i0, i1, 2, 2, C1 * (S(qx, i0, i1, 1, 1) + S(qx, i0, i1, 2, 2)) + C2 * (S(qout, i0, i1, 1, 1) + S(qxx, i0, i1, 3, 3)))
    qxx = put(qxx, i0, i1, npy - 1, npy - 1, C1 * (S(qx, i0, i1, npy - 2, npy - 2) + S(qx, i0, i1, npy - 1, npy - 1)) +
              C2 * (S(qout, i0, i1, npy, npy) + S(qxx, i0, i1, npy - 2, npy - 2)))
    qyy = torch.zeros_like(qin)
    j0, j1, i0, i1 = 2, npy - 1, 3, npx - 2
    qyy = put(qyy, i0, i1, j0, j1, A2 * (S(qy, i0 - 2, i1 - 2, j0, j1) + S(qy, i0 + 1, i1 + 1, j0, j1)) +
              A1 * (S(qy, i0 - 1, i1 - 1, j0, j1) + S(qy, i0, i1, j0, j1)))
    qyy = put(qyy, 2, 2, j0, j1, C1 * (S(qy, 1, 1, j0, j1) + S(qy, 2, 2, j0, j1)) + C2 * (S(qout, 1, 1, j0, j1) + S(qyy, 3, 3, j0, j1)))
    qyy = put(qyy, npx - 1, npx - 1, j0, j1, C1 * (S(qy, npx - 2, npx - 2, j0, j1) + S(qy, npx - 1, npx - 1, j0, j1)) +
              C2 * (S(qout, npx, npx, j0, j1) + S(qyy, npx - 2, npx - 2, j0, j1)))
    qout = put(qout, 2, npx - 1, 2, npy - 1, 0.5 * (S(qxx, 2, npx - 1, 2, npy - 1) + S(qyy, 2, npx - 1, 2, npy - 1)))
    return qout


def a2b_ord2(qin, g):
    """a2b_ord2 (model/a2b_edge_nlm.F90:677-798; TL A2B_ORD2_TLM in model_tlmadm/a2b_edge_tlm.F90): four-cell mean in the interior, edge
    interpolation of the two-cell means across a cube edge, three-cell mean at the cube vertices.  Whole tile; valid on 1..npx, 1..npy."""
    npx, npy = g.npx, g.npy

    def S(a, i0, i1, j0, j1): return a[..., R(j0, j1), R(i0, i1)]
    def Pq(a, i, j): return a[..., j + O, i + O]
    qout = torch.zeros_like(qin)
    i0, i1, j0, j1 = 2, npx - 1, 2, npy - 1
    qout[..., R(j0, j1), R(i0, i1)] = 0.25 * (S(qin, i0 - 1, i1 - 1, j0 - 1, j1 - 1) + S(qin, i0, i1, j0 - 1, j1 - 1) + S(qin, i0 - 1, i1 - 1, j0, j1) + S(qin, i0, i1, j0, j1))
    qout[..., 1 + O, 1 + O] = R3 * (Pq(qin, 1, 1) + Pq(qin, 1, 0) + Pq(qin, 0, 1))
    qout[..., 1 + O, npx + O] = R3 * (Pq(qin, npx - 1, 1) + Pq(qin, npx - 1, 0) + Pq(qin, npx, 1))
    qout[..., npy + O, npx + O] = R3 * (Pq(qin, npx - 1, npy - 1) + Pq(qin, npx, npy - 1) + Pq(qin, npx - 1, npy))
    qout[..., npy + O, 1 + O] = R3 * (Pq(qin, 1, npy - 1) + Pq(qin, 0, npy - 1) + Pq(qin, 1, npy))
    def col(i): return 0.5 * (qin[..., R(1, npy - 1), i - 1 + O] + qin[..., R(1, npy - 1), i + O])     # q2(j), j = 1..npy-1
    def row(j): return 0.5 * (qin[..., j - 1 + O, R(1, npx - 1)] + qin[..., j + O, R(1, npx - 1)])     # q1(i), i = 1..npx-1
    ew = g.edge_w[..., R(2, npy - 1)]; ee = g.edge_e[..., R(2, npy - 1)]
    es = g.edge_s[..., R(2, npx - 1)]; en = g.edge_n[..., R(2, npx - 1)]
    q2 = col(1); qout[..., R(2, npy - 1), 1 + O] = ew * q2[..., :-1] + (1. - ew) * q2[..., 1:]
    q2 = col(npx); qout[..., R(2, npy - 1), npx + O] = ee * q2[..., :-1] + (1. - ee) * q2[..., 1:]
    q1 = row(1); qout[..., 1 + O, R(2, npx - 1)] = es * q1[..., :-1] + (1. - es) * q1[..., 1:]
    q1 = row(npy); qout[..., npy + O, R(2, npx - 1)] = en * q1[..., :-1] + (1. - en) * q1[..., 1:]
    return qout
