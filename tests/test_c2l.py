"""cubed_to_latlon (c2l_ord = 4, model/fv_grid_utils_nlm.F90:2334-2472): the A-grid lon / lat winds fv3jedi_lm's step_nl returns in
traj%ua, traj%va (fv3jedi_lm_dynamics_mod.F90:839-840).  The oracle is pinned physically (solid-body rotation comes back as
(U cos(lat), 0), converging with resolution), the C ABI is compared with the oracle (module c2l_ord4 and fv3lm_step_nl_winds)."""
import numpy as np
import pytest
import torch
from oracle import c2l as oc
from oracle.cubed_sphere import R
from oracle.dyn_core import halo_of
from synth import grid as G
from common import metrics, ograd, handle, rnd, check_module


def zonal_dgrid(N):
    """D-grid winds of the flow (u_east, v_north) = (cos(lat), 0): projections on the cell-edge directions at the edge mid-points"""
    M = metrics(N)
    xyz = G.ll2xyz(M["grid"])
    def edge_wind(p_a, p_b):
        mid = G.mid3(p_a, p_b)
        e = p_b - p_a; e = e / np.linalg.norm(e, axis=-1, keepdims=True)
        ll = G.xyz2ll(mid)
        east = np.stack([-np.sin(ll[..., 0]), np.cos(ll[..., 0]), np.zeros_like(ll[..., 0])], -1)
        return (east * e).sum(-1) * np.cos(ll[..., 1])
    NX = N + 7
    u = np.zeros((6, 1, NX, NX)); v = np.zeros((6, 1, NX, NX))
    u[:, 0, R(1, N + 1), R(1, N)] = edge_wind(xyz[:, R(1, N + 1), R(1, N)], xyz[:, R(1, N + 1), R(2, N + 1)])
    v[:, 0, R(1, N), R(1, N + 1)] = edge_wind(xyz[:, R(1, N), R(1, N + 1)], xyz[:, R(2, N + 1), R(1, N + 1)])
    return u, v, M


def _solid_body_error(N):
    u, v, M = zonal_dgrid(N)
    halo, getb = halo_of(N)
    ut, vt = halo.dgrid(torch.from_numpy(u), torch.from_numpy(v))
    ua, va = oc.c2l_ord4(ut, vt, ograd(N))
    C = (slice(None), 0, R(1, N), R(1, N))
    lat = M["agrid"][:, R(1, N), R(1, N), 1]
    return float(np.abs(ua[C].numpy() - np.cos(lat)).max()), float(np.abs(va[C].numpy()).max())


def test_oracle_solid_body_rotation():
    e12, e24 = _solid_body_error(12), _solid_body_error(24)
    print(e12, e24)
    assert max(e24) < 5e-3                                  # second-order edge rows dominate
    assert e24[0] < 0.5 * e12[0] and e24[1] < 0.5 * e12[1]  # and it converges


def _run(emu, N=12, K=3, seed=5):
    rng = np.random.default_rng(seed)
    M = metrics(N)
    g = ograd(N)
    halo, getb = halo_of(N)
    f = dict(u=10.0 * rnd(rng, N, K), v=10.0 * rnd(rng, N, K))
    for n in ("a11", "a12", "a21", "a22"):
        f[n] = np.ascontiguousarray(M[n][:, None])
    act = ["u", "v"]
    def fn(u, v):
        u, v = halo.dgrid(u, v)
        return oc.c2l_ord4(u, v, g)
    C = (1, N, 1, N)
    return check_module(handle(N, K, emu), "c2l_ord4", N, K, f, act, dict(ua=C, va=C), fn, {}, rng, tol=1e-13, dot_tol=1e-13)


def test_c2l_emu():
    print(_run(True))


def test_c2l_c24_emu():
    print(_run(True, N=24, K=2))


@pytest.mark.gpu
def test_c2l_gpu():
    print(_run(False))


def _step_winds(emu, nonhydro, **kw):
    """fv3lm_step_nl with fv3lm_set_c2l: ua, va of the propagated trajectory vs the oracle's step + cubed_to_latlon"""
    import test_step_api as tsa
    from oracle import fv_dynamics as ofv
    N, K = 12, 4
    h, f, comp, rng, cfg, ak, bk = tsa.make(emu, N, K, nonhydro=nonhydro, extra=kw or None)
    M = metrics(N)
    Cc = (slice(None), R(1, N), R(1, N))
    h.set_c2l(*[h.scatter_c(np.ascontiguousarray(M[n][Cc])) for n in ("a11", "a12", "a21", "a22")])
    # (test_step_api.make feeds whole-cube arrays: hand this handle its own sub-domains)
    h.set_phis(h.scatter_c(np.ascontiguousarray(f["phis"][:, 0][Cc])))
    h.traj_set(0, {k: h.scatter_c(comp[k]) for k in tsa.ACT})
    h.step_nl(0, 1)
    ua = np.zeros((h.nsub, K, h.nyl, h.nxl)); va = np.zeros_like(ua)
    h.traj_get_winds(ua, va)
    ua = h.gather_c(ua, np.zeros((6, K, N, N))); va = h.gather_c(va, np.zeros((6, K, N, N)))
    full = {k: torch.from_numpy(f[k]) for k in tsa.ACT}
    o = ofv.step_nl(full, ograd(N), ak, bk, cfg, torch.from_numpy(f["phis"]), winds=True)
    C = (slice(None), slice(None), R(1, N), R(1, N))
    res = {}
    for nm, a in (("ua", ua), ("va", va)):
        ref = o[nm][C].numpy()
        res[nm] = np.abs(a - ref).max() / np.abs(ref).max()
        assert res[nm] < 2e-9, res
    # the winds are those of the result: close to a plain average of the D-grid winds the same call stored in slot 1
    out = {k: np.zeros((h.nsub, K, h.nyl, h.nxl)) for k in tsa.ACT}
    h.traj_get(1, out)
    assert np.abs(ua).max() > 1.0 and np.abs(ua).max() < 3.0 * np.abs(out["u"]).max() + 3.0 * np.abs(out["v"]).max()
    return res


def test_winds_need_set_c2l_emu():
    import fv3lm
    from test_fv_dynamics import eta
    ak, bk = eta(3, 100.0)
    h = fv3lm.FV3LM(fv3lm.default_config(12, 3), ak, bk, emu=True)
    with pytest.raises(RuntimeError, match="set_c2l"):
        h.traj_get_winds(np.zeros((6, 3, 12, 12)), np.zeros((6, 3, 12, 12)))


def test_step_nl_winds_emu():
    print(_step_winds(True, True))


def test_step_nl_winds_hydro_layout_emu():
    print(_step_winds(True, False, layout_x=2, layout_y=2))


@pytest.mark.gpu
def test_step_nl_winds_gpu():
    print(_step_winds(False, True))
