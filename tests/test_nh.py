"""Non-hydrostatic kernels: Riem_Solver_c / Riem_Solver3 (SIM1), update_dz_c / update_dz_d and the
non-hydrostatic acoustic loop: C-ABI library vs the torch oracle."""
import numpy as np
import pytest
import torch
from oracle import nh as onh
from oracle.dyn_core import halo_of
from common import metrics, ograd, handle, rnd, check_module
from test_dyn_core import CFG, smooth, two_sided_params, TWO_SIDED
from test_fv_dynamics import eta

RD = 8314.47 / 28.965
NHCFG = dict(CFG); NHCFG.update(rdgas=RD, grav=9.80665, p_fac=0.05)


def column_state(N, K, seed, eta_levels=None):
    rng = np.random.default_rng(seed)
    ptop = CFG["ptop"]
    ak, bk = eta(K, ptop)
    if eta_levels is not None:          # (ak, bk) given by the caller, e.g. the bench's 72 levels with ptop = 1 Pa
        ak, bk = eta_levels
        ptop = float(ak[0])
    ps = 1.0e5 + 300.0 * rnd(rng, N, 1)
    pe = ak[None, :, None, None] + bk[None, :, None, None] * ps
    delp = (pe[:, 1:] - pe[:, :-1]) * (1.0 + 0.02 * rnd(rng, N, K))
    T = 280.0 + 3.0 * rnd(rng, N, K)
    pec = np.concatenate([np.full_like(ps, ptop), ptop + np.cumsum(delp, axis=1)], axis=1)
    dz = -(RD * T / 9.80665) * (np.log(pec[:, 1:]) - np.log(pec[:, :-1])) * (1.0 + 0.01 * rnd(rng, N, K))
    zb = 100.0 * rnd(rng, N, 1)
    z = np.concatenate([zb - np.cumsum(dz[:, ::-1], axis=1)[:, ::-1], zb], axis=1)
    pm = 0.5 * (pec[:, 1:] + pec[:, :-1])
    pt = T                                              # the solver takes pt with pkz folded in: any positive field works
    w = 0.5 * rnd(rng, N, K)
    ws = 0.05 * rnd(rng, N, 1)
    return dict(delp=delp, pt=pt, z=z, w=w, ws=ws, zb=zb), rng, ak, bk


def _run_riem(emu, mode, a_imp=1.0, K=5, eta_levels=None):
    N = 12
    f, rng, ak, bk = column_state(N, K, 3 + mode, eta_levels)
    dts = 150.0
    g = ograd(N)
    cfg = dict(NHCFG); cfg["a_imp"] = a_imp
    if eta_levels is not None:
        cfg["ptop"] = float(ak[0])
    if mode == 0:
        f["z"] = f["z"] * cfg["grav"]; f["zb"] = f["zb"] * cfg["grav"]   # Riem_Solver_c works on gz and phis
        def fn(delp, pt, z, w, ws, zb):
            pef, gz = onh.riem_solver_c(dts, delp, pt, z, w, ws, zb, cfg)
            return pef, gz
        onames = ["pp", "z_n"]
        nk = dict(pp=K + 1, z_n=K + 1)
    else:
        def fn(delp, pt, z, w, ws, zb):
            wn, dzn, zh, ppe = onh.riem_solver3(dts, delp, pt, z, w, ws, zb, cfg)
            return ppe, zh, wn, dzn
        onames = ["pp", "z_n", "w_n", "dz_n"]
        nk = dict(pp=K + 1, z_n=K + 1, w_n=K, dz_n=K)
    C = (0, N + 1, 0, N + 1) if mode == 0 else (1, N, 1, N)
    outs = {o: C for o in onames}
    h = handle(N, K, emu, ak, bk, ptop=cfg["ptop"]) if eta_levels is not None else handle(N, K, emu, ak, bk)
    p = dict(mode=mode, dts=dts, ptop=cfg["ptop"], akap=cfg["akap"], rdgas=cfg["rdgas"], grav=cfg["grav"], a_imp=a_imp)
    return check_module(h, "riem", N, K, f, list(f.keys()), outs, fn, p, rng, tol=1e-10, dot_tol=1e-13, pert_scale=1e-3, out_nk=nk)     # achieved 6.7e-12 (SIM solver) / 1e-15


def _run_dzc(emu):
    N, K = 12, 4
    rng = np.random.default_rng(9)
    ak, bk = eta(K, CFG["ptop"])
    g = ograd(N)
    area = metrics(N)["area"][:, None]
    f = dict(ut=0.1 * np.abs(area) * rnd(rng, N, K), vt=0.1 * np.abs(area) * rnd(rng, N, K))
    zs = 100.0 * rnd(rng, N, 1)
    dz = 800.0 * (1.0 + 0.1 * rnd(rng, N, K))
    f["gz"] = np.concatenate([zs + np.cumsum(dz[:, ::-1], axis=1)[:, ::-1], zs], axis=1)
    f["zs"] = zs
    dp0 = [(ak[k + 1] - ak[k]) + (bk[k + 1] - bk[k]) * 1.e5 for k in range(K)]
    dts = 150.0
    def fn(ut, vt, gz):
        return onh.update_dz_c(dts, dp0, torch.from_numpy(zs), ut, vt, gz, g)
    h = handle(N, K, emu, ak, bk)
    C = (0, N + 1, 0, N + 1)
    return check_module(h, "update_dz_c", N, K, f, ["ut", "vt", "gz"], {"gz_n": C, "ws": C}, fn, dict(dts=dts), rng, tol=1e-11, dot_tol=1e-12,
                        out_nk=dict(gz_n=K + 1, ws=1))


def _run_dzd(emu):
    N, K = 12, 4
    rng = np.random.default_rng(10)
    ak, bk = eta(K, CFG["ptop"])
    g = ograd(N)
    area = metrics(N)["area"][:, None]
    f = dict(crx=0.2 * rnd(rng, N, K), cry=0.2 * rnd(rng, N, K), xfx=0.1 * np.abs(area) * rnd(rng, N, K), yfx=0.1 * np.abs(area) * rnd(rng, N, K))
    zs = 100.0 * rnd(rng, N, 1)
    dz = 800.0 * (1.0 + 0.1 * rnd(rng, N, K))
    f["zh"] = np.concatenate([zs + np.cumsum(dz[:, ::-1], axis=1)[:, ::-1], zs], axis=1)
    f["zs"] = zs
    dp0 = [(ak[k + 1] - ak[k]) + (bk[k + 1] - bk[k]) * 1.e5 for k in range(K)]
    cfg = dict(CFG)
    prm = __import__("oracle.dyn_core", fromlist=["x"]).level_params(cfg, K)
    dts = 300.0
    def fn(crx, cry, xfx, yfx, zh):
        return onh.update_dz_d(prm["nord_v"], prm["damp_v"], 2, dp0, torch.from_numpy(zs), zh, crx, cry, xfx, yfx, g, 1.0 / dts)
    h = handle(N, K, emu, ak, bk)
    C = (1, N, 1, N)
    p = dict(cfg); p.update(dts=dts, do_vort_damp=1)
    return check_module(h, "update_dz_d", N, K, f, ["crx", "cry", "xfx", "yfx", "zh"], {"zh_n": C, "ws": C}, fn, p, rng, tol=1e-11, dot_tol=1e-12,
                        out_nk=dict(zh_n=K + 1, ws=1))


def nh_state(N, K, seed, ak, bk):
    rng = np.random.default_rng(seed)
    halo, getb = halo_of(N)
    ps = 1.0e5 + 100.0 * smooth(rng, N, 1)
    pe = ak[None, :, None, None] + bk[None, :, None, None] * ps
    delp = pe[:, 1:] - pe[:, :-1]
    pm = 0.5 * (pe[:, 1:] + pe[:, :-1])
    T = 280.0 + 1.0 * smooth(rng, N, K)
    delz = -(RD * T / 9.80665) * (np.log(pe[:, 1:]) - np.log(pe[:, :-1]))
    pkz = np.exp(CFG["akap"] * np.log(-RD / 9.80665 * delp * T / delz))
    f = dict(u=5.0 * smooth(rng, N, K), v=5.0 * smooth(rng, N, K), pt=T / pkz, delp=delp, w=0.05 * smooth(rng, N, K), delz=delz,
             phis=50.0 * 9.80665 * smooth(rng, N, 1))
    t = {k: torch.from_numpy(a) for k, a in f.items()}
    for k in ("delp", "pt", "phis"):
        t[k] = halo.scalar(t[k])
    t["u"], t["v"] = getb(t["u"], t["v"]); t["u"], t["v"] = halo.dgrid(t["u"], t["v"])
    return {k: a.numpy().copy() for k, a in t.items()}, rng


def _run_dyn_nh(emu, n_split, a_imp=1.0, extra=None):
    N, K = 12, 4
    ak, bk = eta(K, CFG["ptop"])
    f, rng = nh_state(N, K, 17, ak, bk)
    g = ograd(N)
    cfg = dict(NHCFG); cfg.update(n_split=n_split, bdt=600.0, a_imp=a_imp)
    cfg.update(extra or {})
    act = ["u", "v", "pt", "delp", "w", "delz"]
    onames = ["u_n", "v_n", "pt_n", "delp_n", "w_n", "delz_n", "mfx", "cx"]
    key = dict(u_n="u", v_n="v", pt_n="pt", delp_n="delp", w_n="w", delz_n="delz")
    def fn(*a):
        st = {n: torch.from_numpy(f[n]) for n in f}
        st.update(dict(zip(act, a)))
        o = onh.dyn_core_nh(st, g, cfg, ak, bk)
        return tuple(o[key.get(k, k)] for k in onames)
    C = (1, N, 1, N); npx = N + 1
    outs = dict(u_n=(1, N, 1, npx), v_n=(1, npx, 1, N), pt_n=C, delp_n=C, w_n=C, delz_n=C, mfx=(1, npx, 1, N), cx=(1, npx, -2, N + 3))
    h = handle(N, K, emu, ak, bk)
    p = two_sided_params(cfg); p.update(do_vort_damp=int(cfg["do_vort_damp"]), hydrostatic=0)
    return check_module(h, "dyn_core_nh", N, K, f, act, outs, fn, p, rng, tol=5e-11, dot_tol=1e-12, pert_scale=1e-3)     # achieved 2.4e-12 / 1e-15 on the B200


@pytest.mark.parametrize("mode", [0, 1])
def test_riem_emu(mode):
    print(_run_riem(True, mode))


@pytest.mark.parametrize("a_imp", [0.75, 0.999])
def test_riem3_sim_solver_emu(a_imp):
    """Riem_Solver3 with the off-centred SIM_solver (reference default a_imp = 0.75, model/fv_arrays_nlm.F90:357)"""
    print(_run_riem(True, 1, a_imp))


def test_riem_c_keeps_sim1_emu():
    """Riem_Solver_c calls SIM1_solver for every a_imp > 0.5 (model/nh_utils_nlm.F90:366-376)"""
    print(_run_riem(True, 0, 0.75))


def test_dyn_core_nh_sim_solver_emu():
    print(_run_dyn_nh(True, 2, 0.75))


def test_dyn_core_nh_two_sided_emu():
    """two sets of switches in the non-hydrostatic loop (adds the w transport and the zh transport of update_dz_d, nh_utils_tlm.F90:496)"""
    print(_run_dyn_nh(True, 2, extra=TWO_SIDED))


@pytest.mark.gpu
def test_dyn_core_nh_two_sided_gpu():
    _run_dyn_nh(False, 2, extra=TWO_SIDED)


def test_dyn_core_nh_heat_emu():
    """d_con = 1 in the non-hydrostatic loop: w-damping and KE-damping heat, pkz from delz (dyn_core_nlm.F90:1078-1099)"""
    print(_run_dyn_nh(True, 2, extra=dict(d_con=1.0)))


def test_dyn_core_nh_beta_emu():
    """beta = 0.4: split_p_grad (model/dyn_core_nlm.F90:1531-1643, TL dyn_core_tlm.F90:3592-3757) replaces nh_p_grad; three acoustic
    sub-steps so that the blended hydrostatic gradient du / dv is handed on twice (beta_d = 0 on the first, :373-375)"""
    print(_run_dyn_nh(True, 3, extra=dict(beta=0.4)))


@pytest.mark.gpu
def test_dyn_core_nh_beta_gpu():
    _run_dyn_nh(False, 3, extra=dict(beta=0.4))


def test_update_dz_c_emu():
    print(_run_dzc(True))


def test_update_dz_d_emu():
    print(_run_dzd(True))


def test_dyn_core_nh_emu():
    print(_run_dyn_nh(True, 2))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1])
def test_riem_gpu(mode):
    _run_riem(False, mode)


@pytest.mark.gpu
def test_dyn_core_nh_heat_gpu():
    _run_dyn_nh(False, 2, extra=dict(d_con=1.0))


@pytest.mark.gpu
def test_riem3_sim_solver_gpu():
    _run_riem(False, 1, 0.75)
    _run_dyn_nh(False, 2, 0.75)


@pytest.mark.gpu
def test_update_dz_gpu():
    _run_dzc(False); _run_dzd(False)


@pytest.mark.gpu
def test_dyn_core_nh_gpu():
    _run_dyn_nh(False, 2)
