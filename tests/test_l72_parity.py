"""L72 parity against the oracle on the workload bench.py's cpu_baseline arm pushes through it: C12, 72 levels with the bench's eta
levels (ptop = 1 Pa, thin top layers), non-hydrostatic, dt = 6750 s, n_split = 7 (bench.py: cpu_baseline / model_config).

Every other oracle comparison in the suite uses 2..10 levels; this one exercises the multi-layer search of the vertical remap
(model_tlmadm/fv_mapz_tlm.F90:7909-8046 MAP1_PPM_TLM), the 72-level Thomas recurrences of the semi-implicit solver
(model_tlmadm/nh_utils_tlm.F90:2548 SIM1_SOLVER_TLM) and cs_profile (fv_mapz_tlm.F90:8513) at full depth:
  * one whole model step NL / TL (oracle jvp) / AD (oracle vjp), element-wise, + the dot-product identity;
  * the `remap` and `riem` modules at K = 72 on columns of that state.
The achieved errors are recorded (tests/common.py: record -> profiles/parity_errors_*.json); the tolerances below are about 10x the
values achieved on the B200 and in the host emulation.
"""
import numpy as np
import pytest
import torch
import fv3lm
import common
from common import metrics, ograd, handle, rnd, check_module
from oracle import fv_dynamics as ofv
from oracle.cubed_sphere import R
from synth import state as S

ZVIR = (8314.47 / 18.015) / (8314.47 / 28.965) - 1.0
RD = 8314.47 / 28.965
N, K = 12, 72
ACT = ["u", "v", "t", "delp", "qv", "ql", "qi", "o3", "w", "delz"]


def bench_workload():
    """the sample of bench.py: cpu_baseline (same generator, same seeds, same switches)"""
    dt = 450.0 * 180.0 / N
    n_split = S.n_split_auto(N, dt, False)
    ak, bk = S.eta_levels(K)
    M = metrics(N)
    st = S.make_state(M, K, ak, bk, hydrostatic=False)
    pert = S.make_pert(st, 1)
    yvec = S.make_pert(st, 2)
    mc = dict(n_split=n_split, k_split=1, dt=dt, ptop=1.0, d2_bg_k1=0.20, d2_bg_k2=0.10, hydrostatic=0, zvir=ZVIR)
    ocfg = dict(nord=1, d2_bg=0.015, d2_bg_k1=0.2, d2_bg_k2=0.1, n_sponge=0, vtdm4=0.0005, do_vort_damp=True, dddmp=0.2, d4_bg=0.15,
                hord_mt=2, hord_vt=2, hord_tm=2, hord_dp=2, hord_tr=2, n_sponge_ord=0, ptop=1.0, akap=2.0 / 7.0, cp_air=3.5 * RD,
                zvir=ZVIR, hydrostatic=False, k_split=1, n_split=n_split, dt=dt, rdgas=RD, grav=9.80665, p_fac=0.05)
    return ak, bk, st, pert, yvec, mc, ocfg


def full(a):
    NX = N + 7
    z = np.zeros(a.shape[:-2] + (NX, NX)); z[..., R(1, N), R(1, N)] = a
    return z


_oracle_cache = {}


def oracle_step():
    """NL, TL = jvp, AD = vjp of the oracle's step on the bench sample (tens of seconds of CPU work; shared by the emu and gpu tests)"""
    if _oracle_cache:
        return _oracle_cache
    ak, bk, st, pert, yvec, mc, ocfg = bench_workload()
    g = ograd(N)
    x = tuple(torch.from_numpy(full(st[k])) for k in ACT)
    phis = torch.from_numpy(full(st["phis"][:, None]))
    def fn(*a):
        o = ofv.step_nl(dict(zip(ACT, a)), g, ak, bk, ocfg, phis)
        return tuple(o[k] for k in ACT)
    dx = tuple(torch.from_numpy(full(pert[k])) for k in ACT)
    yb = tuple(torch.from_numpy(full(yvec[k])) for k in ACT)
    nl, tl = torch.func.jvp(fn, x, dx)
    _, vjp = torch.func.vjp(fn, *x)
    ad = vjp(yb)
    C = (slice(None), slice(None), R(1, N), R(1, N))
    _oracle_cache.update(nl={k: a[C].numpy() for k, a in zip(ACT, nl)}, tl={k: a[C].numpy() for k, a in zip(ACT, tl)},
                         ad={k: a[C].numpy() for k, a in zip(ACT, ad)})
    return _oracle_cache


# achieved (B200 / host emulation): see profiles/parity_errors_gpu.json, parity_errors_emu.json
TOL_NL, TOL_TL, TOL_AD, TOL_DOT = 5e-11, 2e-11, 2e-12, 1e-13


def _step(emu):
    ak, bk, st, pert, yvec, mc, ocfg = bench_workload()
    ref = oracle_step()
    h = handle(N, K, emu, ak, bk, **mc)
    h.set_phis(st["phis"])
    h.traj_set(0, {k: st[k] for k in ACT})
    h.step_nl(0, 1)
    out = {k: np.zeros_like(st[k]) for k in ACT}
    h.traj_get(1, out)
    errs = {}
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
    for k in ACT:
        errs["nl." + k] = rel(out[k], ref["nl"][k])
    mdx = {k: pert[k].copy() for k in ACT}
    h.step_tl(0, mdx)
    for k in ACT:
        errs["tl." + k] = rel(mdx[k], ref["tl"][k])
    mty = {k: yvec[k].copy() for k in ACT}
    h.step_ad(0, mty)
    for k in ACT:
        errs["ad." + k] = rel(mty[k], ref["ad"][k])
    lhs = sum((mdx[k] * yvec[k]).sum() for k in ACT)
    rhs = sum((pert[k] * mty[k]).sum() for k in ACT)
    errs["dot"] = abs(lhs - rhs) / max(abs(lhs), abs(rhs))
    common.record(errs)
    for k, e in errs.items():
        tol = TOL_DOT if k == "dot" else TOL_NL if k.startswith("nl.") else TOL_TL if k.startswith("tl.") else TOL_AD
        assert e <= tol, (k, e, tol)
    return errs


def test_step_l72_emu():
    print(_step(True))


@pytest.mark.gpu
def test_step_l72_gpu():
    print(_step(False))


# ---- modules at K = 72 with the bench's eta levels -----------------------------------------------------------------------------
def _remap72(emu, last_step):
    """Lagrangian_to_Eulerian on 72-level columns: Lagrangian surfaces displaced by up to 8 % of the layer thickness from the
    reference ones, so that target layers straddle one, two or three source layers (the data-dependent search of map1_ppm)"""
    from oracle import fv_mapz as omap
    rng = np.random.default_rng(72)
    g = ograd(N)
    ak, bk = S.eta_levels(K)
    ptop = 1.0
    akap = 2.0 / 7.0
    ps = 1.0e5 + 500.0 * rnd(rng, N, 1)
    pe_ref = ak[None, :, None, None] + bk[None, :, None, None] * ps
    dref = pe_ref[:, 1:] - pe_ref[:, :-1]
    dlag = dref * (1.0 + 0.08 * rnd(rng, N, K))
    pe = np.concatenate([np.full_like(ps, ptop), ptop + np.cumsum(dlag, axis=1)], axis=1)
    f = dict(pe=pe, pk=np.exp(akap * np.log(pe)), peln=np.log(pe), pt=300.0 + 10.0 * rnd(rng, N, K),
             q0=0.01 * (1.0 + 0.3 * rnd(rng, N, K)), u=10.0 * rnd(rng, N, K), v=10.0 * rnd(rng, N, K))
    cfg = dict(ptop=ptop, akap=akap, cp_air=3.5 * RD, zvir=ZVIR, hydrostatic=True)
    act = list(f.keys())
    def fn(*a):
        d = dict(zip(act, a))
        st = dict(pe=d["pe"], pk=d["pk"], peln=d["peln"], pt=d["pt"], q=[d["q0"]], u=d["u"], v=d["v"],
                  delp=torch.zeros_like(d["pt"]), pkz=torch.zeros_like(d["pt"]))
        o = omap.lagrangian_to_eulerian(st, g, ak, bk, cfg, last_step)
        return (o["pt"], o["q"][0], o["u"], o["v"], o["delp"], o["pkz"])
    C = (1, N, 1, N)
    outs = dict(pt_n=C, q0_n=C, u_n=(1, N, 1, N + 1), v_n=(1, N + 1, 1, N), delp_n=C, pkz_n=C)
    h = handle(N, K, emu, ak, bk, ptop=ptop)
    p = dict(ptop=ptop, akap=akap, zvir=ZVIR, last_step=int(last_step))
    return check_module(h, "remap", N, K, f, act, outs, fn, p, rng, tol=1e-10, dot_tol=1e-12, pert_scale=1e-3)


@pytest.mark.parametrize("last_step", [True, False])
def test_remap_l72_emu(last_step):
    print(_remap72(True, last_step))


@pytest.mark.gpu
@pytest.mark.parametrize("last_step", [True, False])
def test_remap_l72_gpu(last_step):
    print(_remap72(False, last_step))


def _riem72(emu, mode):
    """Riem_Solver_c (mode 0) / Riem_Solver3 (mode 1) on the 72-level columns of the bench sample"""
    import test_nh
    return test_nh._run_riem(emu, mode, K=K, eta_levels=S.eta_levels(K))


@pytest.mark.parametrize("mode", [0, 1])
def test_riem_l72_emu(mode):
    print(_riem72(True, mode))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1])
def test_riem_l72_gpu(mode):
    print(_riem72(False, mode))
