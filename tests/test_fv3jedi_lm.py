"""The host-side mirror of the reference's top level (src/fv3jedi_lm_mod.F90) used the way FV3-JEDI uses the reference: create,
propagate the trajectory over a window with step_nl, run the tangent-linear model forward and the adjoint backward over the saved
trajectories, and check <M dx, y> = <dx, M^T y> for the whole window, dynamics + turbulence."""
import numpy as np
import pytest
from fv3jedi_lm import fv3jedi_lm_type
from oracle.cubed_sphere import R
from common import metrics
from test_dyn_core import CFG
from test_fv_dynamics import eta, api_state, ZVIR
import test_turbulence as tt

N, K = 12, 4


def bl_driver(traj):
    """stand-in for BL_DRIVER: trajectory-dependent, diagonally dominant systems (stronger mixing where delp is larger)"""
    rng = np.random.default_rng(11)
    co = tt.coeffs(rng, traj["delp"].shape)
    w = traj["delp"] / traj["delp"].mean()
    for s in "vsq":
        co["ak" + s] = co["ak" + s] * w; co["ck" + s] = co["ck" + s] * w
        co["bk" + s] = 1.0 - co["ak" + s] - co["ck" + s] + 0.05
    return co


def make(do_dyn=1, do_phy_trb=1, nonhydro=True, **flags):
    ptop = CFG["ptop"]
    ak, bk = eta(K, ptop)
    lm = fv3jedi_lm_type(flags=dict(hydrostatic=0 if nonhydro else 1, n_split=1, k_split=1, kappa=CFG["akap"], cp=CFG["cp_air"], zvir=ZVIR,
                                    d2_bg_k1=CFG["d2_bg_k1"], d2_bg_k2=CFG["d2_bg_k2"], **flags),
                         metrics=metrics(N), bl_driver=bl_driver, emu=True)
    lm.conf.do_dyn = do_dyn; lm.conf.do_phy_trb = do_phy_trb; lm.conf.do_phy_mst = 0
    lm.create(900.0, N + 1, N + 1, K, ptop, ak, bk)
    f, rng = api_state(N, K, 41, ak, bk, nonhydro)
    C = (slice(None), slice(None), R(1, N), R(1, N))
    for k in lm._names:
        lm.traj[k][...] = f[k][C]
    lm.traj["phis"][...] = f["phis"][:, 0][:, R(1, N), R(1, N)]
    return lm, rng


def test_window_dot_product_emu():
    lm, rng = make()
    names = lm._names
    nt = 2
    lm.conf.nt = nt
    # trajectory over the window (FV3-JEDI stores it per time level)
    trajs = []
    lm.init_nl()
    for n in range(1, nt + 1):
        lm.conf.n = n
        trajs.append({k: v.copy() for k, v in lm.traj.items()})
        lm.step_nl()
    lm.final_nl()
    assert np.abs(lm.traj["ua"]).max() > 0.0                      # a11 .. a22 were in the metrics: the winds came back
    assert all(np.isfinite(lm.traj[k]).all() for k in names)
    dx = {k: rng.standard_normal(lm.pert[k].shape) * (np.abs(trajs[0][k]).mean() * 1e-3 + 1e-12) for k in names}
    y = {k: rng.standard_normal(lm.pert[k].shape) / (np.abs(trajs[0][k]).mean() + 1e-30) for k in names}
    # tangent linear, forward
    lm.init_tl()
    for k in names:
        lm.pert[k][...] = dx[k]
    for n in range(1, nt + 1):
        lm.conf.n = n
        for k in trajs[n - 1]:
            lm.traj[k][...] = trajs[n - 1][k]
        lm.step_tl()
    lm.final_tl()
    mdx = {k: lm.pert[k].copy() for k in names}
    assert not lm.pert["ua"].any() and not lm.pert["va"].any()     # internal part of pert stays zero (:243-252)
    # adjoint, backward
    lm.init_ad()
    for k in names:
        lm.pert[k][...] = y[k]
    for n in range(nt, 0, -1):
        lm.conf.n = n
        for k in trajs[n - 1]:
            lm.traj[k][...] = trajs[n - 1][k]
        lm.step_ad()
    lm.final_ad()
    lhs = sum((mdx[k] * y[k]).sum() for k in names)
    rhs = sum((dx[k] * lm.pert[k]).sum() for k in names)
    assert abs(lhs - rhs) <= 1e-10 * max(abs(lhs), abs(rhs)), (lhs, rhs)
    lm.delete()


def test_switches_emu():
    """do_dyn = 0: physics only; do_phy_trb = 0 (and mst = 0): do_phy falls to 0 at create (:85) and no bl_driver is needed"""
    lm, rng = make(do_dyn=0)
    dx = {k: rng.standard_normal(lm.pert[k].shape) for k in lm._names}
    for k in lm._names:
        lm.pert[k][...] = dx[k]
    lm.step_tl()
    from oracle import turbulence as otb
    lt = otb.set_ltraj(bl_driver(lm.traj), lm.traj["delp"], lm.conf.ptop, CFG["akap"])
    ref = otb.step(lt, dx, CFG["akap"], 1)
    for k in tt.FLD:
        assert np.abs(lm.pert[k] - ref[k]).max() <= 1e-12 * np.abs(ref[k]).max(), k
    assert all(np.array_equal(lm.pert[k], dx[k]) for k in ("delp", "w", "delz"))
    lm2, _ = make(do_phy_trb=0, nonhydro=False)
    assert lm2.conf.do_phy == 0
    lm2.bl_driver = None
    lm2.pert["u"][...] = 1.0
    lm2.step_tl()                                                   # dynamics only
    assert np.isfinite(lm2.pert["u"]).all()


def test_errors_emu():
    ak, bk = eta(K, CFG["ptop"])
    lm = fv3jedi_lm_type(emu=True)                                  # conf defaults like the reference: do_phy_mst = 1
    with pytest.raises(RuntimeError, match="moist"):
        lm.create(900.0, N + 1, N + 1, K, CFG["ptop"], ak, bk)
    with pytest.raises(ValueError):
        lm.create(900.0, N + 1, N + 1, K, CFG["ptop"], ak[:-1], bk)
    lm3, _ = make()
    lm3.bl_driver = None
    with pytest.raises(RuntimeError, match="bl_driver"):
        lm3.step_tl()


@pytest.mark.gpu
def test_window_gpu():
    """one TL / AD pair through the mirror on the device"""
    ptop = CFG["ptop"]
    ak, bk = eta(K, ptop)
    lm = fv3jedi_lm_type(flags=dict(hydrostatic=0, n_split=1, kappa=CFG["akap"], cp=CFG["cp_air"], zvir=ZVIR, d2_bg_k1=CFG["d2_bg_k1"],
                                    d2_bg_k2=CFG["d2_bg_k2"]), metrics=metrics(N), bl_driver=bl_driver, emu=False)
    lm.conf.do_phy_mst = 0
    lm.create(900.0, N + 1, N + 1, K, ptop, ak, bk)
    f, rng = api_state(N, K, 41, ak, bk, True)
    C = (slice(None), slice(None), R(1, N), R(1, N))
    for k in lm._names:
        lm.traj[k][...] = f[k][C]
    lm.traj["phis"][...] = f["phis"][:, 0][:, R(1, N), R(1, N)]
    dx = {k: rng.standard_normal(lm.pert[k].shape) * (np.abs(lm.traj[k]).mean() * 1e-3 + 1e-12) for k in lm._names}
    y = {k: rng.standard_normal(lm.pert[k].shape) for k in lm._names}
    for k in lm._names:
        lm.pert[k][...] = dx[k]
    lm.step_tl()
    mdx = {k: lm.pert[k].copy() for k in lm._names}
    for k in lm._names:
        lm.pert[k][...] = y[k]
    lm.step_ad()
    lhs = sum((mdx[k] * y[k]).sum() for k in lm._names)
    rhs = sum((dx[k] * lm.pert[k]).sum() for k in lm._names)
    assert abs(lhs - rhs) <= 1e-10 * max(abs(lhs), abs(rhs)), (lhs, rhs)
