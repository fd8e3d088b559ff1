"""fv_tp_2d: C-ABI library vs the torch oracle (NL, TL = jvp, AD = vjp)."""
import numpy as np
import pytest
import torch
import fv3lm
from oracle import tp_core as otp
from common import metrics, ograd, handle, rnd, relerr, region

TOL = 1e-12   # fp64, summation-order differences only


def _inputs(N, K, seed):
    rng = np.random.default_rng(seed)
    M = metrics(N)
    area = M["area"][:, None]
    f = {}
    f["q"] = rnd(rng, N, K, 1.0, 10.0)
    f["crx"] = rnd(rng, N, K, 0.3)
    f["cry"] = rnd(rng, N, K, 0.3)
    f["xfx"] = rnd(rng, N, K, 0.2) * np.abs(area) * 0.5
    f["yfx"] = rnd(rng, N, K, 0.2) * np.abs(area) * 0.5
    f["ra_x"] = np.abs(area) * (1.0 + 0.1 * rnd(rng, N, K))
    f["ra_y"] = np.abs(area) * (1.0 + 0.1 * rnd(rng, N, K))
    return f, rng


def _oracle(f, N, hord):
    g = ograd(N)
    names = ["q", "crx", "cry", "xfx", "yfx", "ra_x", "ra_y"]
    def fn(*a):
        d = dict(zip(names, a))
        fx, fy, q = otp.fv_tp_2d(d["q"], d["crx"], d["cry"], hord, d["xfx"], d["yfx"], g, d["ra_x"], d["ra_y"])
        return fx, fy
    return names, fn


def _run(emu, hord):
    N, K = 12, 3
    f, rng = _inputs(N, K, 1234)
    names, fn = _oracle(f, N, hord)
    tin = [torch.from_numpy(f[n].copy()) for n in names]
    h = handle(N, K, emu)
    out = {"fx": np.zeros_like(f["q"]), "fy": np.zeros_like(f["q"])}
    # ---- NL
    traj = {n: f[n].copy() for n in names}; traj.update({k: v.copy() for k, v in out.items()})
    h.module_run("fv_tp_2d", fv3lm.MODE_NL, traj, params={"hord": hord})
    fx_o, fy_o = fn(*tin)
    assert relerr(region(traj["fx"], 1, N + 1, 1, N), region(fx_o.numpy(), 1, N + 1, 1, N)) < TOL
    assert relerr(region(traj["fy"], 1, N, 1, N + 1), region(fy_o.numpy(), 1, N, 1, N + 1)) < TOL
    # ---- TL
    dp = {n: rnd(rng, N, K) * (np.abs(f[n]).mean() * 1e-2) for n in names}
    traj = {n: f[n].copy() for n in names}; traj.update({k: v.copy() for k, v in out.items()})
    pert = {n: dp[n].copy() for n in names}; pert.update({k: v.copy() for k, v in out.items()})
    h.module_run("fv_tp_2d", fv3lm.MODE_TL, traj, pert, params={"hord": hord})
    _, (dfx, dfy) = torch.func.jvp(fn, tuple(tin), tuple(torch.from_numpy(dp[n]) for n in names))
    assert relerr(region(pert["fx"], 1, N + 1, 1, N), region(dfx.numpy(), 1, N + 1, 1, N)) < TOL
    assert relerr(region(pert["fy"], 1, N, 1, N + 1), region(dfy.numpy(), 1, N, 1, N + 1)) < TOL
    tl_fx, tl_fy = pert["fx"].copy(), pert["fy"].copy()
    # ---- AD
    yfx_ = np.zeros_like(f["q"]); yfy_ = np.zeros_like(f["q"])
    region(yfx_, 1, N + 1, 1, N)[...] = region(rnd(rng, N, K), 1, N + 1, 1, N)
    region(yfy_, 1, N, 1, N + 1)[...] = region(rnd(rng, N, K), 1, N, 1, N + 1)
    traj = {n: f[n].copy() for n in names}; traj.update({k: v.copy() for k, v in out.items()})
    pert = {n: np.zeros_like(f[n]) for n in names}; pert["fx"] = yfx_.copy(); pert["fy"] = yfy_.copy()
    h.module_run("fv_tp_2d", fv3lm.MODE_AD, traj, pert, params={"hord": hord})
    _, vjp = torch.func.vjp(fn, *tin)
    ad_o = vjp((torch.from_numpy(yfx_), torch.from_numpy(yfy_)))
    for n, a in zip(names, ad_o):
        assert relerr(pert[n], a.numpy()) < TOL, n
    # ---- dot-product test  <M dx, y> = <dx, M^T y>
    lhs = (tl_fx * yfx_).sum() + (tl_fy * yfy_).sum()
    rhs = sum((dp[n] * pert[n]).sum() for n in names)
    assert abs(lhs - rhs) <= 1e-13 * max(abs(lhs), abs(rhs))


def _run_mono(emu, hord, seed=1234, positive=True):
    """monotone PPM schemes of the nonlinear model (hord 8..13, tp_core_nlm.F90:470-578): trajectory values only"""
    N, K = 12, 3
    f, rng = _inputs(N, K, seed)
    if not positive:
        f["q"] = f["q"] - 10.0          # sign changes: exercises the positive-definite constraint of hord 9 / 13
    f["q"][:, 0] = np.round(f["q"][:, 0])   # plateaus: dm = 0, bl * br >= 0 branches
    names, fn = _oracle(f, N, hord)
    tin = [torch.from_numpy(f[n].copy()) for n in names]
    h = handle(N, K, emu)
    traj = {n: f[n].copy() for n in names}; traj.update(fx=np.zeros_like(f["q"]), fy=np.zeros_like(f["q"]))
    h.module_run("fv_tp_2d", fv3lm.MODE_NL, traj, params={"hord": hord})
    fx_o, fy_o = fn(*tin)
    assert relerr(region(traj["fx"], 1, N + 1, 1, N), region(fx_o.numpy(), 1, N + 1, 1, N)) < TOL
    assert relerr(region(traj["fy"], 1, N, 1, N + 1), region(fy_o.numpy(), 1, N, 1, N + 1)) < TOL


@pytest.mark.parametrize("hord", [3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13])
def test_fv_tp_2d_monotone_emu(hord):
    _run_mono(True, hord)
    _run_mono(True, hord, seed=99, positive=False)


@pytest.mark.gpu
def test_fv_tp_2d_monotone_gpu():
    for hord in (8, 9, 10, 13):
        _run_mono(False, hord)
        _run_mono(False, hord, seed=99, positive=False)


@pytest.mark.parametrize("hord", [1, 2, 333])
def test_fv_tp_2d_emu(hord):
    _run(True, hord)


@pytest.mark.gpu
@pytest.mark.parametrize("hord", [1, 2, 333])
def test_fv_tp_2d_gpu(hord):
    _run(False, hord)
