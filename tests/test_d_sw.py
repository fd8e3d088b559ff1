"""d_sw (D-grid step) and a2b_ord4: C-ABI library vs the torch oracle."""
import numpy as np
import pytest
import torch
from oracle import d_sw as odsw
from oracle.a2b_edge import a2b_ord4
from common import metrics, ograd, handle, rnd, check_module, region
import fv3lm


def dsw_inputs(N, K, seed):
    rng = np.random.default_rng(seed)
    f = {}
    f["delp"] = 1000.0 + 50.0 * rnd(rng, N, K)
    f["pt"] = 300.0 + 5.0 * rnd(rng, N, K)
    f["u"] = 10.0 * rnd(rng, N, K)
    f["v"] = 10.0 * rnd(rng, N, K)
    f["w"] = 0.5 * rnd(rng, N, K)
    f["uc"] = 10.0 * rnd(rng, N, K)
    f["vc"] = 10.0 * rnd(rng, N, K)
    f["ua"] = 10.0 * rnd(rng, N, K)
    f["va"] = 10.0 * rnd(rng, N, K)
    f["divg_d"] = 1e-5 * rnd(rng, N, K)
    return f, rng


def level_params(K, sponge, hord=2):
    """per-level switches like dyn_core_nlm.F90:579-625 with a 1-layer sponge at k=0"""
    p = dict(hord_mt=[hord] * K, hord_vt=[hord] * K, hord_tm=[hord] * K, hord_dp=[hord] * K, nord=[1] * K, nord_v=[1] * K,
             nord_w=[1] * K, nord_t=[1] * K, d2_bg=[0.015] * K, damp_v=[0.0005] * K, damp_w=[0.0005] * K, damp_t=[0.0005] * K)
    if sponge:
        for n in ("hord_mt", "hord_vt", "hord_tm", "hord_dp"):
            p[n][0] = 1
        p["nord"][0] = 0; p["d2_bg"][0] = 4.0
        p["nord_w"][0] = 0; p["damp_w"][0] = 4.0
        p["nord_v"][0] = 0; p["damp_v"][0] = 2.0
        p["nord_t"][0] = 0; p["damp_t"][0] = 0.0005
    return p


def flat_params(p, K):
    out = {}
    for name, v in p.items():
        if isinstance(v, list):
            out[name] = v[-1]
            for k in range(K):
                if v[k] != v[-1]:
                    out["%s@%d" % (name, k)] = v[k]
        else:
            out[name] = v
    return out


def _run_dsw(emu, hydrostatic, sponge, hord=2, d_con=0.0, pert=None, traj_over=None):
    """pert: perturbation-side overrides of level_params (split_hord / split_damp of the TL/AD model)"""
    N, K = 12, 2
    f, rng = dsw_inputs(N, K, 11)
    g = ograd(N)
    dt = 450.0
    prm = level_params(K, sponge, hord)
    prm.update(dddmp=0.2, d4_bg=0.15, hydrostatic=hydrostatic)
    if d_con > 0.0:
        prm["d_con"] = [0.0 if (sponge and k == 0) else d_con for k in range(K)]     # d_con_k = 0 in sponge layers (dyn_core_nlm.F90:595-628)
    prm.update(traj_over or {})
    pp = None
    if pert is not None:
        pp = dict(prm); pp.update(pert)
    names = list(f.keys())
    act = [n for n in names if not (hydrostatic and n == "w")]
    onames = ["delp_n", "pt_n", "u_n", "v_n", "fx", "fy", "crx", "cry", "xfx", "yfx"] + ([] if hydrostatic else ["w_n"]) + (["heat"] if d_con > 0.0 else [])
    key = dict(delp_n="delp", pt_n="pt", u_n="u", v_n="v", w_n="w")
    def fn(*a):
        d = {n: torch.from_numpy(f[n]) for n in names}
        d.update(dict(zip(act, a)))
        o = odsw.d_sw(d["delp"], d["pt"], d["u"], d["v"], d["w"], d["uc"], d["vc"], d["ua"], d["va"], d["divg_d"], g, dt, prm, pp)
        return tuple(o[key.get(k, k)] for k in onames)
    npx = N + 1
    C = (1, N, 1, N)
    outs = dict(delp_n=C, pt_n=C, w_n=C, heat=C, u_n=(1, N, 1, npx), v_n=(1, npx, 1, N), fx=(1, npx, 1, N), fy=(1, N, 1, npx),
                crx=(1, npx, -2, N + 3), xfx=(1, npx, -2, N + 3), cry=(-2, N + 3, 1, npx), yfx=(-2, N + 3, 1, npx))
    outs = {k: outs[k] for k in onames}
    h = handle(N, K, emu)
    p = flat_params(prm, K); p["dt"] = dt; p["hydrostatic"] = int(hydrostatic)
    if pert is not None:
        for k_, v_ in flat_params({k: pp[k] for k in pert}, K).items():
            p[k_ if k_ == "split_damp" else "p." + k_] = v_
    return check_module(h, "d_sw", N, K, f, act, outs, fn, p, rng, tol=2e-12, dot_tol=1e-12)


def _run_a2b(emu, N=12):
    K = 2
    rng = np.random.default_rng(3)
    f = {"qin": rnd(rng, N, K)}
    g = ograd(N)
    h = handle(N, K, emu)
    return check_module(h, "a2b_ord4", N, K, f, ["qin"], {"qout": (1, N + 1, 1, N + 1)}, lambda q: (a2b_ord4(q, g),), {}, rng)


def _a2b_nl_ad(emu, N):
    """NL result and the adjoint of a random output adjoint through the module interface (used to compare implementations)"""
    K = 2
    rng = np.random.default_rng(31)
    q = rnd(rng, N, K)
    y = np.zeros_like(q)
    region(y, 1, N + 1, 1, N + 1)[...] = region(rnd(rng, N, K), 1, N + 1, 1, N + 1)
    h = handle(N, K, emu)
    traj = {"qin": q.copy(), "qout": np.zeros_like(q)}
    h.module_run("a2b_ord4", fv3lm.MODE_NL, traj)
    nl = region(traj["qout"], 1, N + 1, 1, N + 1).copy()
    traj = {"qin": q.copy(), "qout": np.zeros_like(q)}
    pert = {"qin": np.zeros_like(q), "qout": y.copy()}
    h.module_run("a2b_ord4", fv3lm.MODE_AD, traj, pert)
    return nl, pert["qin"].copy()


def test_a2b_ord4_emu():
    _run_a2b(True)


@pytest.mark.parametrize("hydrostatic,sponge", [(True, False), (False, True)])
def test_d_sw_emu(hydrostatic, sponge):
    _run_dsw(True, hydrostatic, sponge)


@pytest.mark.parametrize("hydrostatic,sponge", [(True, False), (False, True)])
def test_d_sw_heat_emu(hydrostatic, sponge):
    """d_con > 0: kinetic energy lost to the damping returned as heat_s (sw_core_nlm.F90:1494-1525)"""
    _run_dsw(True, hydrostatic, sponge, d_con=1.0)


@pytest.mark.gpu
def test_d_sw_heat_gpu():
    _run_dsw(False, False, True, d_con=1.0)


SPLIT_HORD = dict(hord_mt=[1, 333], hord_vt=[1, 333], hord_tm=[1, 1], hord_dp=[1, 333])
SPLIT_DAMP = dict(split_damp=1, nord=[0, 0], d2_bg=[0.05, 0.02], nord_v=[0, 1], damp_v=[0.03, 0.001], nord_t=[0, 0], damp_t=[0.02, 0.001], dddmp=0.1, d4_bg=0.12)


def test_d_sw_split_hord_emu():
    """split_hord: the perturbation is transported with its own (different) linear scheme, linearised about the same inputs, while
    the trajectory keeps the nonlinear model's scheme (model_tlmadm/sw_core_tlm.F90:1664-1682, 1987-1997)"""
    _run_dsw(True, False, False, pert=SPLIT_HORD)


def test_d_sw_split_damp_emu():
    """split_damp: divergence / vorticity / tracer damping with the perturbation-side coefficients (sw_core_tlm.F90:2341-2366, 2436-2451)"""
    _run_dsw(True, False, True, pert=SPLIT_DAMP)
    both = dict(SPLIT_HORD); both.update(SPLIT_DAMP)
    _run_dsw(True, True, False, pert=both)


LIN2 = dict(hord_mt=[2, 2], hord_vt=[2, 2], hord_tm=[2, 2], hord_dp=[2, 2])


@pytest.mark.parametrize("hord", [3, 4, 5, 6, 7, 8, 9, 10])
def test_d_sw_monotone_trajectory_emu(hord):
    """operational-like split: the trajectory runs the nonlinear model's monotone PPM (hord 8 / 9 / 10 in fv_tp_2d and xtp_u / ytp_v),
    the perturbation the linear scheme hord_pert = 2 linearised about it (fv_arrays_tlmadm.F90:40-46)"""
    _run_dsw(True, False, True, hord=hord, pert=LIN2)


@pytest.mark.gpu
def test_d_sw_monotone_trajectory_gpu():
    _run_dsw(False, False, True, hord=10, pert=LIN2)
    _run_dsw(False, True, False, hord=9, pert=LIN2)


def test_d_sw_split_damp_vort_damp_one_side_emu():
    """the reference's defaults: do_vort_damp = F for the trajectory, do_vort_damp_pert = T for the increment
    (model/fv_arrays_nlm.F90 vs fv_arrays_tlmadm.F90:37-92): the diffusive flux is added on one side only (sw_core_tlm.F90:2506-2530)"""
    N, K = 12, 2
    pert = dict(SPLIT_DAMP)
    _run_dsw(True, False, False, pert=pert, traj_over=dict(damp_v=[0.0, 0.0], damp_t=[0.0, 0.0], damp_w=[0.0, 0.0]))
    _run_dsw(True, False, False, pert=dict(split_damp=1, damp_v=[0.0, 0.0]))        # and the other way round


@pytest.mark.gpu
def test_d_sw_split_gpu():
    both = dict(SPLIT_HORD); both.update(SPLIT_DAMP)
    _run_dsw(False, False, True, pert=both)


def test_d_sw_hord333_emu():
    """third-order linear scheme (hord = 333) in fv_tp_2d and xtp_u / ytp_v, first-order sponge layer on top"""
    _run_dsw(True, False, True, 333)


@pytest.mark.gpu
def test_d_sw_hord333_gpu():
    _run_dsw(False, False, True, 333)


@pytest.mark.gpu
def test_a2b_ord4_gpu():
    _run_a2b(False)


@pytest.mark.gpu
@pytest.mark.parametrize("hydrostatic,sponge", [(True, False), (False, True), (False, False)])
def test_d_sw_gpu(hydrostatic, sponge):
    _run_dsw(False, hydrostatic, sponge)
