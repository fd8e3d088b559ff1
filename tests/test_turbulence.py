"""Linearised boundary-layer turbulence (src/physics/turbulence/fv3jedi_lm_turbulence_mod.F90 step_nl :149,
step_tl :218, step_ad :285, vtrilupert :562, vtrisolvepert :583): the oracle pinned against linear algebra, then the C ABI
(fv3lm_turb_*) against the oracle for TL, AD and NL, the dot-product test, the device-resident variants chained with the dynamics
step the way fv3jedi_lm_mod does (src/fv3jedi_lm_mod.F90:165-185), sub-domain layouts and the error paths."""
import numpy as np
import pytest
import fv3lm
from oracle import turbulence as otb
from common import handle
from test_dyn_core import CFG
from test_fv_dynamics import eta, ZVIR

FLD = ["u", "v", "t", "qv", "qi", "ql", "o3"]
KAPPA = CFG["akap"]


def coeffs(rng, shape):
    """diagonally dominant systems of the kind an implicit diffusion step produces: a = -alpha(k), c = -alpha(k+1), b = 1 - a - c
    (+ a little extra on the diagonal so that the surface row is not singular)"""
    co = {}
    for s in "vsq":
        al = 0.05 + 2.0 * rng.random(shape)
        a = -al.copy(); a[..., 0, :, :] = 0.0
        c = np.zeros(shape); c[..., :-1, :, :] = -al[..., 1:, :, :]
        co["ak" + s] = a; co["ck" + s] = c
        co["bk" + s] = 1.0 - a - c + 0.1 * rng.random(shape)
    return co


def state(rng, shape, scale=1.0):
    return {n: scale * rng.standard_normal(shape) for n in FLD + ["delp", "w", "delz"]}


def delp_of(rng, shape, ptop):
    K = shape[-3]
    ak, bk = eta(K, ptop)
    ps = 1.0e5 + 500.0 * rng.standard_normal(shape[:-3] + (1,) + shape[-2:])
    pe = ak[:, None, None] + bk[:, None, None] * ps
    return pe[..., 1:, :, :] - pe[..., :-1, :, :]


# ---- the oracle itself --------------------------------------------------------------------------------------------------------
def test_oracle_solves_the_tridiagonal_system():
    """ygswitch = 1: (vtrilupert, vtrisolvepert phase 1) is x = A^-1 y for A = tridiag(a, b, c)"""
    rng = np.random.default_rng(1)
    shape = (9, 2, 3)
    co = coeffs(rng, shape)
    a, b = otb.vtrilupert(co["akv"], co["bkv"], co["ckv"])
    y = rng.standard_normal(shape)
    x = otb.vtrisolvepert(a, b, co["ckv"], y, 1, 1)
    K = shape[0]
    for j in range(shape[1]):
        for i in range(shape[2]):
            A = np.diag(co["bkv"][:, j, i]) + np.diag(co["akv"][1:, j, i], -1) + np.diag(co["ckv"][:-1, j, i], 1)
            np.testing.assert_allclose(x[:, j, i], np.linalg.solve(A, y[:, j, i]), rtol=1e-11, atol=1e-13)
    assert K == 9


@pytest.mark.parametrize("yg", [1, 0])
def test_oracle_adjoint_is_the_transpose(yg):
    """phase 2 applied to unit vectors gives the transposed matrix of phase 1, for both surface treatments"""
    rng = np.random.default_rng(2)
    K = 7
    shape = (K, 1, 1)
    co = coeffs(rng, shape)
    a, b = otb.vtrilupert(co["akq"], co["bkq"], co["ckq"])
    E = np.eye(K).reshape(K, K, 1, 1)
    M1 = np.stack([otb.vtrisolvepert(a, b, co["ckq"], E[n], 1, yg)[:, 0, 0] for n in range(K)], axis=1)
    M2 = np.stack([otb.vtrisolvepert(a, b, co["ckq"], E[n], 2, yg)[:, 0, 0] for n in range(K)], axis=1)
    np.testing.assert_allclose(M2, M1.T, rtol=1e-12, atol=1e-14)


# ---- C ABI vs oracle ----------------------------------------------------------------------------------------------------------
def make(emu, N=12, K=6, seed=3, **kw):
    ptop = CFG["ptop"]
    ak, bk = eta(K, ptop)
    h = handle(N, K, emu, ak, bk, ptop=ptop, kappa=KAPPA, zvir=ZVIR, **kw)
    rng = np.random.default_rng(seed)
    shape = (6, K, N, N)
    co = coeffs(rng, shape)
    traj = state(rng, shape)
    traj["delp"] = delp_of(rng, shape, ptop)
    traj["t"] = 280.0 + 5.0 * traj["t"]
    lt = otb.set_ltraj(co, traj["delp"], ptop, KAPPA)
    return h, rng, shape, co, traj, lt


def golden_inputs(N=12, K=3, seed=9):
    """seeded inputs of the committed fixture tests/golden/physics_c12.npz (tools/make_golden.py)"""
    rng = np.random.default_rng(seed)
    shape = (6, K, N, N)
    co = coeffs(rng, shape)
    traj = state(rng, shape)
    traj["delp"] = delp_of(rng, shape, CFG["ptop"])
    return shape, co, traj, state(rng, shape), state(rng, shape)


def up(h, d):
    """whole-cube compute-domain arrays -> this handle's sub-domains (contiguous)"""
    return {k: h.scatter_c(v).copy() for k, v in d.items()}


def down(h, d, shape):
    return {k: h.gather_c(v, np.zeros(shape)) for k, v in d.items()}


def err(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _run(emu, **kw):
    h, rng, shape, co, traj, lt = make(emu, **kw)
    h.traj_set(0, up(h, traj))
    h.turb_set_ltraj(0, up(h, co))                       # raw diagonals; pk from the delp of slot 0; decomposition on the device
    res = {}
    # TL
    dx = state(rng, shape)
    a = up(h, dx)
    h.turb_step_tl(0, a)
    a = down(h, a, shape)
    ref = otb.step(lt, dx, KAPPA, 1)
    for n in FLD:
        res["tl." + n] = err(a[n], ref[n]); assert res["tl." + n] < 1e-12, ("tl", n, res)
    assert all(np.array_equal(a[n], dx[n]) for n in ("delp", "w", "delz"))
    # AD
    y = state(rng, shape)
    b = up(h, y)
    h.turb_step_ad(0, b)
    b = down(h, b, shape)
    ref = otb.step(lt, y, KAPPA, 2)
    for n in FLD:
        res["ad." + n] = err(b[n], ref[n]); assert res["ad." + n] < 1e-12, ("ad", n, res)
    # dot-product test across the ABI
    lhs = sum((a[n] * y[n]).sum() for n in FLD)
    rhs = sum((dx[n] * b[n]).sum() for n in FLD)
    res["dot"] = abs(lhs - rhs) / max(abs(lhs), abs(rhs))
    assert res["dot"] < 1e-13, res
    # NL: the same solves on the trajectory fields of another slot (the state after the dynamics step), in place
    tr2 = {k: (v if k == "delp" else 1.7 * v) for k, v in traj.items()}
    h.traj_set(2, up(h, tr2))
    h.turb_step_nl(0, 2)
    out = {n: np.zeros_like(v) for n, v in up(h, traj).items()}
    h.traj_get(2, out)
    out = down(h, out, shape)
    ref = otb.step(lt, tr2, KAPPA, 1)
    for n in FLD:
        res["nl." + n] = err(out[n], ref[n]); assert res["nl." + n] < 1e-12, ("nl", n, res)
    # coefficients decomposed by the caller + pk given: same result
    h.traj_set(1, up(h, traj))
    h.turb_set_ltraj(1, up(h, lt), decomposed=True)
    c = up(h, dx)
    h.turb_step_tl(1, c)
    c = down(h, c, shape)
    for n in FLD:
        assert err(c[n], a[n]) < 1e-14, n
    return res


def test_turbulence_emu():
    print(_run(True))


def test_turbulence_nonhydro_emu():
    """w and delz are part of the increment but not touched by the scheme"""
    print(_run(True, hydrostatic=0))


def test_turbulence_layout_2x2_emu():
    print(_run(True, layout_x=2, layout_y=2))


def _chain(emu):
    """fv3jedi_lm_mod step_tl = dynamics then physics (src/fv3jedi_lm_mod.F90:165-170), step_ad = physics then dynamics (:179-185),
    on the device-resident increments: the composite still passes the dot-product test"""
    from test_step_api import make as make_step
    import test_step_api as tsa
    N, K = 12, 4
    h, f, comp, rng, cfg, ak, bk = make_step(emu, N, K, nonhydro=True)
    ACT = tsa.ACT
    shape = (6, K, N, N)
    h.turb_set_ltraj(0, coeffs(rng, shape))
    dx = {k: rng.standard_normal(comp[k].shape) * (np.abs(comp[k]).mean() * 1e-3) for k in ACT}
    y = {k: rng.standard_normal(comp[k].shape) / (np.abs(comp[k]).mean() + 1e-30) for k in ACT}
    a = {k: dx[k].copy() for k in ACT}
    h.pert_upload(a); h.step_tl_dev(0); h.turb_step_tl_dev(0); h.pert_download(a)
    b = {k: y[k].copy() for k in ACT}
    h.pert_upload(b); h.turb_step_ad_dev(0); h.step_ad_dev(0); h.pert_download(b)
    lhs = sum((a[k] * y[k]).sum() for k in ACT)
    rhs = sum((dx[k] * b[k]).sum() for k in ACT)
    assert abs(lhs - rhs) <= 1e-10 * max(abs(lhs), abs(rhs)), (lhs, rhs)
    # and the physics did something
    c = {k: dx[k].copy() for k in ACT}
    h.step_tl(0, c)
    assert np.abs(c["u"] - a["u"]).max() > 0.0 and np.array_equal(c["delp"], a["delp"]) and np.array_equal(c["w"], a["w"])


def test_dynamics_then_turbulence_emu():
    _chain(True)


def test_turbulence_errors_emu():
    h, rng, shape, co, traj, lt = make(True, K=5)
    with pytest.raises(RuntimeError, match="turb_set_ltraj"):
        h.turb_step_tl(7, up(h, state(rng, shape)))
    bad = dict(co); del bad["bkq"]
    with pytest.raises(RuntimeError, match="bkq"):
        h.turb_set_ltraj(0, bad)
    with pytest.raises(RuntimeError, match="slot"):
        h.turb_set_ltraj(5, co)                           # no pk and no trajectory in slot 5


@pytest.mark.gpu
def test_turbulence_gpu():
    print(_run(False))
    print(_run(False, hydrostatic=0, layout_x=2, layout_y=2))


@pytest.mark.gpu
def test_dynamics_then_turbulence_gpu():
    _chain(False)
