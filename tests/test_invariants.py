"""Physical invariants of the nonlinear step: they pin the cube-edge / corner handling
independently of the oracle (which was restated by the same hands as the kernels).  One fv3jedi_lm dynamics step
on a synthetic C12 state through the step-level C ABI:
  * dry-mass conservation      sum(area * (ps - ptop)) is unchanged by transport + remap (flux form, closed sphere)
  * tracer-mass conservation   sum(area * delp * q) is unchanged for every tracer
  * constancy                  a tracer that is uniform stays uniform (delp and q see the same mass fluxes)
  * TL of a conserved quantity the tangent of the total mass is zero for any increment (M dx conserves mass)
"""
import numpy as np
import pytest
import fv3lm
from synth import grid as G, state as S
from oracle.cubed_sphere import R

ZVIR = (8314.47 / 18.015) / (8314.47 / 28.965) - 1.0


def _setup(emu, nonhydro, N=12, K=6):
    dt = 900.0
    ak, bk = S.eta_levels(K, ptop=300.0)
    M = G.build_metrics(N)
    cfg = fv3lm.default_config(N, K, n_split=3, k_split=1, dt=dt, ptop=300.0, d2_bg_k1=0.20, d2_bg_k2=0.10, hydrostatic=0 if nonhydro else 1, zvir=ZVIR)
    h = fv3lm.FV3LM(cfg, ak, bk, emu=emu)
    h.set_metrics(M)
    st = S.make_state(M, K, ak, bk, hydrostatic=not nonhydro)
    st["o3"] = np.full_like(st["o3"], 3.0e-6)            # a uniform tracer
    fields = [f for f in h.FIELDS if f in st]
    h.set_phis(st["phis"])
    h.traj_set(0, {k: st[k] for k in fields})
    area = np.ascontiguousarray(M["area"][:, R(1, N), R(1, N)])
    return h, st, fields, area


def _check(emu, nonhydro):
    h, st, fields, area = _setup(emu, nonhydro)
    h.step_nl(0, 1)
    out = {k: np.zeros_like(st[k]) for k in fields}
    h.traj_get(1, out)
    a = area[:, None]
    m0 = (a * st["delp"]).sum(); m1 = (a * out["delp"]).sum()
    assert abs(m1 - m0) <= 1e-13 * abs(m0), ("dry mass", m0, m1)
    for q in ("qv", "ql", "qi", "o3"):
        t0 = (a * st["delp"] * st[q]).sum(); t1 = (a * out["delp"] * out[q]).sum()
        assert abs(t1 - t0) <= 1e-12 * abs(t0), (q, t0, t1)
    assert np.abs(out["o3"] / 3.0e-6 - 1.0).max() <= 1e-12, "uniform tracer"
    # tangent of the conserved totals
    dx = S.make_pert(st, 5)
    d = {k: np.ascontiguousarray(dx[k]) for k in fields}
    dm0 = (a * d["delp"]).sum(); dsc = (a * np.abs(d["delp"])).sum()
    dq0 = (a * (d["delp"] * st["qv"] + st["delp"] * d["qv"])).sum()
    qsc = (a * (np.abs(d["delp"]) * st["qv"] + st["delp"] * np.abs(d["qv"]))).sum()
    h.step_tl(0, d)
    dm1 = (a * d["delp"]).sum()
    dq1 = (a * (d["delp"] * out["qv"] + out["delp"] * d["qv"])).sum()
    assert abs(dm1 - dm0) <= 1e-12 * dsc, ("TL dry mass", dm0, dm1)
    assert abs(dq1 - dq0) <= 1e-11 * qsc, ("TL vapour mass", dq0, dq1)
    h.close()


@pytest.mark.parametrize("nonhydro", [False, True])
def test_invariants_emu(nonhydro):
    _check(True, nonhydro)


@pytest.mark.gpu
@pytest.mark.parametrize("nonhydro", [False, True])
def test_invariants_gpu(nonhydro):
    _check(False, nonhydro)


def test_grid_metrics_invariants():
    """geometry of the synthetic cubed sphere (the reference ships no grid data): positive areas, sum(area) = 4 pi R^2,
    the 8 cube vertices have identical dual areas, halo metrics equal the owner tile's values across an aligned edge"""
    N = 12
    M = G.build_metrics(N)
    o = 2
    c = (slice(None), R(1, N), R(1, N))
    assert (M["area"][c] > 0).all() and (M["area_c"][:, R(1, N + 1), R(1, N + 1)] > 0).all()
    tot = M["area"][c].sum()
    assert abs(tot / (4.0 * np.pi * 6371.0e3 ** 2) - 1.0) < 1e-12
    ac = M["area_c"]
    corners = np.array([[ac[t, 1 + o, 1 + o], ac[t, 1 + o, N + 1 + o], ac[t, N + 1 + o, N + 1 + o], ac[t, N + 1 + o, 1 + o]] for t in range(6)])
    assert np.abs(corners / corners[0, 0] - 1.0).max() < 1e-12
    for name in ("dxa", "dya", "area", "dxc", "dyc", "area_c"):
        a = M[name]
        # tile 1 east <-> tile 2 west (aligned contact 1): halo column of one = first interior column of the other
        assert np.abs(a[0, R(1, N), N + 1 + o] - a[1, R(1, N), 1 + o]).max() <= 1e-12 * np.abs(a[1, R(1, N), 1 + o]).max(), name
