"""ORACLE (test infrastructure only).  cubed_to_latlon with c2l_ord = 4 restated in torch float64 from
model/fv_grid_utils_nlm.F90:2334-2472 (c2l_ord4; non-nested, grid_type < 4), called with mode = 1 at the end of fv_dynamics
(model/fv_dynamics_nlm.F90:738); a11 .. a22 from init_cubed_to_latlon :2248-2310 (synth/grid.py).
Written the way the reference is: interior first, then the four edge blocks in its order (later blocks overwrite the corners).
parity unpinned (no reference vectors); pinned physically by solid-body rotation (tests/test_c2l.py)."""
import torch
from .sw_core import S, Z, put

A1, A2, C1, C2 = 0.5625, -0.0625, 1.125, -0.125


def c2l_ord4(u, v, g):
    """u, v: D-grid winds [6, K, NY, NX] with valid halos.  Returns ua, va (compute domain filled, zero elsewhere)."""
    N = g.N
    npx = npy = N + 1
    is_, ie, js, je = 1, N, 1, N
    dx, dy = g.dx, g.dy
    utmp = Z(u); vtmp = Z(u)
    i0, i1, j0, j1 = max(2, is_), min(npx - 2, ie), max(2, js), min(npy - 2, je)
    utmp = put(utmp, i0, i1, j0, j1, C2 * (S(u, i0, i1, j0 - 1, j1 - 1) + S(u, i0, i1, j0 + 2, j1 + 2)) + C1 * (S(u, i0, i1, j0, j1) + S(u, i0, i1, j0 + 1, j1 + 1)))
    vtmp = put(vtmp, i0, i1, j0, j1, C2 * (S(v, i0 - 1, i1 - 1, j0, j1) + S(v, i0 + 2, i1 + 2, j0, j1)) + C1 * (S(v, i0, i1, j0, j1) + S(v, i0 + 1, i1 + 1, j0, j1)))

    def edge(utmp, vtmp, i0, i1, j0, j1):
        wu0 = S(u, i0, i1, j0, j1) * S(dx, i0, i1, j0, j1); wu1 = S(u, i0, i1, j0 + 1, j1 + 1) * S(dx, i0, i1, j0 + 1, j1 + 1)
        wv0 = S(v, i0, i1, j0, j1) * S(dy, i0, i1, j0, j1); wv1 = S(v, i0 + 1, i1 + 1, j0, j1) * S(dy, i0 + 1, i1 + 1, j0, j1)
        utmp = put(utmp, i0, i1, j0, j1, 2. * (wu0 + wu1) / (S(dx, i0, i1, j0, j1) + S(dx, i0, i1, j0 + 1, j1 + 1)))
        vtmp = put(vtmp, i0, i1, j0, j1, 2. * (wv0 + wv1) / (S(dy, i0, i1, j0, j1) + S(dy, i0 + 1, i1 + 1, j0, j1)))
        return utmp, vtmp
    utmp, vtmp = edge(utmp, vtmp, is_, ie, 1, 1)                   # js == 1        :2392-2403
    utmp, vtmp = edge(utmp, vtmp, is_, ie, npy - 1, npy - 1)       # je + 1 == npy  :2405-2417
    utmp, vtmp = edge(utmp, vtmp, 1, 1, js, je)                    # is == 1        :2419-2434
    utmp, vtmp = edge(utmp, vtmp, npx - 1, npx - 1, js, je)        # ie + 1 == npx  :2436-2449
    ut = S(utmp, is_, ie, js, je); vt = S(vtmp, is_, ie, js, je)
    ua = put(Z(u), is_, ie, js, je, S(g.a11, is_, ie, js, je) * ut + S(g.a12, is_, ie, js, je) * vt)     # :2453-2458
    va = put(Z(u), is_, ie, js, je, S(g.a21, is_, ie, js, je) * ut + S(g.a22, is_, ie, js, je) * vt)
    return ua, va
