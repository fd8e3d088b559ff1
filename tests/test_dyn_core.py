"""dyn_core (acoustic sub-cycle incl. halo exchanges): C-ABI library vs the torch oracle."""
import numpy as np
import pytest
import torch
from oracle import dyn_core as odyn
from oracle.dyn_core import halo_of
from common import metrics, ograd, handle, rnd, check_module

RD = 8314.47 / 28.965
CFG = dict(nord=1, d2_bg=0.015, d2_bg_k1=0.2, d2_bg_k2=0.1, n_sponge=0, vtdm4=0.0005, do_vort_damp=True, dddmp=0.2, d4_bg=0.15,
           hord_mt=2, hord_vt=2, hord_tm=2, hord_dp=2, n_sponge_ord=0, ptop=100.0, akap=2.0 / 7.0, cp_air=3.5 * RD)


def smooth(rng, N, K, nsm=4):
    """white noise smoothed a little so that transported fields stay physical"""
    a = rnd(rng, N, K)
    for _ in range(nsm):
        a[..., 1:-1, 1:-1] = 0.2 * (a[..., 1:-1, 1:-1] + a[..., :-2, 1:-1] + a[..., 2:, 1:-1] + a[..., 1:-1, :-2] + a[..., 1:-1, 2:])
    return a


def hydro_state(N, K, seed):
    """a resting, nearly isothermal atmosphere + noise; halos filled consistently"""
    rng = np.random.default_rng(seed)
    halo, getb = halo_of(N)
    ps = 1.0e5 + 100.0 * smooth(rng, N, 1)
    ptop = CFG["ptop"]
    bk = np.linspace(0.0, 1.0, K + 1); ak = ptop * (1.0 - bk)
    pe = ak[None, :, None, None] + bk[None, :, None, None] * ps
    delp = pe[:, 1:] - pe[:, :-1]
    pm = 0.5 * (pe[:, 1:] + pe[:, :-1])
    T = 280.0 + 1.0 * smooth(rng, N, K)
    pt = T / (pm / 1.0e5) ** CFG["akap"]               # potential temperature: cp*pt*dpk is the geopotential increment
    u = 5.0 * smooth(rng, N, K); v = 5.0 * smooth(rng, N, K)
    phis = 50.0 * 9.80665 * smooth(rng, N, 1)
    f = dict(u=u, v=v, pt=pt, delp=delp, w=np.zeros_like(u), phis=phis)
    t = {k: torch.from_numpy(a) for k, a in f.items()}
    t["delp"] = halo.scalar(t["delp"]); t["pt"] = halo.scalar(t["pt"]); t["phis"] = halo.scalar(t["phis"])
    t["u"], t["v"] = getb(t["u"], t["v"])
    t["u"], t["v"] = halo.dgrid(t["u"], t["v"])
    return {k: a.numpy().copy() for k, a in t.items()}, rng


def two_sided_params(cfg):
    """oracle cfg -> module parameters: in two-sided mode the top-level switches are the perturbation model's and cfg["traj"]
    the nonlinear model's ("t.<name>" parameters)"""
    p = {k: v for k, v in cfg.items() if k != "traj"}
    if cfg.get("traj"):
        p["two_sided"] = 1
        for k, v in cfg["traj"].items():
            p["t." + k] = int(v) if isinstance(v, bool) else v
    return p


# perturbation-side switches like the defaults of fv_flags_pert_type (fv_arrays_tlmadm.F90:37-92) on a coarse test grid, with a
# different set for the nonlinear model: split_hord (333 / 1 vs 2), split_damp (nord, coefficients, sponge depth)
TWO_SIDED = dict(hord_mt=333, hord_vt=1, hord_tm=333, hord_dp=1, hord_tr=1, nord=1, dddmp=0.2, d2_bg=0.015, d4_bg=0.12, vtdm4=0.0008,
                 do_vort_damp=True, d2_bg_k1=0.3, d2_bg_k2=0.2, d2_bg_ks=0.1, n_sponge=3, split_damp=True, hord_ks_pert=True, hord_ks_traj=True,
                 traj=dict(hord_mt=2, hord_vt=2, hord_tm=2, hord_dp=2, hord_tr=2, nord=2, dddmp=0.1, d2_bg=0.01, d4_bg=0.15, vtdm4=0.0005,
                           do_vort_damp=True, d2_bg_k1=0.2, d2_bg_k2=0.1, n_sponge=0))


# operational-like: perturbation flags at the defaults of fv_flags_pert_type (linear hord 2, first-order sponge transport) around a
# trajectory that runs the nonlinear model's default monotone schemes (hord_mt = hord_vt = hord_tm = hord_dp = 9, hord_tr = 12,
# model/fv_arrays_nlm.F90:264-277)
TWO_SIDED_MONO = dict(hord_mt=2, hord_vt=2, hord_tm=2, hord_dp=2, hord_tr=2, nord=1, dddmp=0.2, d2_bg=0.015, d4_bg=0.15, vtdm4=0.0005,
                      do_vort_damp=True, d2_bg_k1=0.3, d2_bg_k2=0.2, d2_bg_ks=0.1, n_sponge=2, split_damp=True, hord_ks_pert=True, hord_ks_traj=True,
                      traj=dict(hord_mt=9, hord_vt=9, hord_tm=9, hord_dp=9, hord_tr=12, nord=1, dddmp=0.0, d2_bg=0.0, d4_bg=0.16, vtdm4=0.0005,
                                do_vort_damp=True, d2_bg_k1=0.2, d2_bg_k2=0.1, n_sponge=1, kord_mt=10, kord_wz=10, kord_tm=9, kord_tr=9))


# the reference's own defaults on both sides: trajectory without vorticity damping and with dddmp = d2_bg = 0,
# perturbation with do_vort_damp_pert = T and a deep first-order sponge (n_sponge_pert = 9 > npz of the test)
REF_DEFAULTS = dict(hord_mt=2, hord_vt=2, hord_tm=2, hord_dp=2, hord_tr=2, nord=1, dddmp=0.2, d2_bg=0.015, d4_bg=0.15, vtdm4=0.0005,
                    do_vort_damp=True, d2_bg_k1=4.0, d2_bg_k2=2.0, d2_bg_ks=2.0, n_sponge=9, split_damp=True, hord_ks_pert=True, hord_ks_traj=True,
                    traj=dict(hord_mt=9, hord_vt=9, hord_tm=9, hord_dp=9, hord_tr=12, nord=1, dddmp=0.0, d2_bg=0.0, d4_bg=0.16, vtdm4=0.0,
                              do_vort_damp=False, d2_bg_k1=0.2, d2_bg_k2=0.1, n_sponge=1, kord_mt=8, kord_wz=8, kord_tm=8, kord_tr=8))


def _run(emu, n_split, K=3, modes=("nl", "tl", "ad"), extra=None):
    N = 12
    f, rng = hydro_state(N, K, 5)
    g = ograd(N)
    cfg = dict(CFG); cfg["n_split"] = n_split; cfg["bdt"] = 900.0
    cfg.update(extra or {})
    act = ["u", "v", "pt", "delp"]
    onames = ["u_n", "v_n", "pt_n", "delp_n", "mfx", "mfy", "cx", "cy", "pkz"]
    key = dict(u_n="u", v_n="v", pt_n="pt", delp_n="delp")
    def fn(*a):
        st = {n: torch.from_numpy(f[n]) for n in f}
        st.update(dict(zip(act, a)))
        o = odyn.dyn_core_hydro(st, g, cfg)
        return tuple(o[key.get(k, k)] for k in onames)
    npx = N + 1
    C = (1, N, 1, N)
    outs = dict(u_n=(1, N, 1, npx), v_n=(1, npx, 1, N), pt_n=C, delp_n=C, mfx=(1, npx, 1, N), mfy=(1, N, 1, npx),
                cx=(1, npx, -2, N + 3), cy=(-2, N + 3, 1, npx), pkz=C)
    h = handle(N, K, emu)
    p = two_sided_params(cfg); p["do_vort_damp"] = int(cfg["do_vort_damp"]); p["hydrostatic"] = 1
    return check_module(h, "dyn_core", N, K, f, act, outs, fn, p, rng, tol=5e-11, dot_tol=1e-11, pert_scale=1e-3)


@pytest.mark.parametrize("n_split", [1, 2])
def test_dyn_core_hydro_emu(n_split):
    r = _run(True, n_split)
    print(r)


def _run_del2_cubed(emu, nmax):
    """del2_cubed (model/dyn_core_nlm.F90:2090-2199): nmax passes of the 5-point filter with corner averaging"""
    N, K = 12, 2
    rng = np.random.default_rng(77)
    halo, _ = odyn.halo_of(N)
    g = ograd(N)
    f = {"q": rnd(rng, N, K)}
    h = handle(N, K, emu)
    def fn(q):
        return (odyn.del2_cubed(q, 0.18 * g.da_min, g, nmax, halo),)
    return check_module(h, "del2_cubed", N, K, f, ["q"], {"q_n": (1, N, 1, N)}, fn, dict(cd=0.18, nmax=nmax), rng, tol=1e-12, dot_tol=1e-12)


@pytest.mark.parametrize("nmax", [1, 2, 3])
def test_del2_cubed_emu(nmax):
    print(_run_del2_cubed(True, nmax))


def _run_heat_update(emu, hydrostatic):
    """the end of dyn_core on its own (dyn_core_nlm.F90:1052-1099) with a heat source large enough to reach the delt_max limiter"""
    N, K = 12, 4
    rng = np.random.default_rng(78)
    halo, _ = odyn.halo_of(N)
    g = ograd(N)
    RD = 8314.47 / 28.965
    cfg = dict(CFG); cfg.update(bdt=2.0e-3, hydrostatic=hydrostatic, rdgas=RD, grav=9.80665)
    delp = 1000.0 + 50.0 * rnd(rng, N, K)
    f = dict(heat=3.0e3 * rnd(rng, N, K), pt=300.0 + 5.0 * rnd(rng, N, K), delp=delp)
    f["aux"] = (1.0 + 0.1 * rnd(rng, N, K)) if hydrostatic else -(800.0 + 20.0 * rnd(rng, N, K))
    def fn(heat, pt, delp, aux):
        return (odyn.heat_update(heat, pt, delp, aux, g, cfg, halo, hydrostatic),)
    h = handle(N, K, emu)
    p = dict(cfg); p["hydrostatic"] = int(hydrostatic); p["do_vort_damp"] = 1
    r = check_module(h, "heat_update", N, K, f, ["heat", "pt", "delp", "aux"], {"pt_n": (1, N, 1, N)}, fn, p, rng, tol=1e-12, dot_tol=1e-12, pert_scale=1e-4)
    # both branches of the limiter are exercised
    dtmp = f["heat"] / ((cfg["cp_air"] if hydrostatic else cfg["cp_air"] - RD) * f["delp"])
    frac = float((np.abs(dtmp) > abs(cfg["bdt"])).mean())
    assert 0.1 < frac < 0.9, frac
    return r


@pytest.mark.parametrize("hydrostatic", [True, False])
def test_heat_update_emu(hydrostatic):
    print(_run_heat_update(True, hydrostatic))


def test_dyn_core_hydro_two_sided_emu():
    """the TL/AD model's two sets of switches: perturbation schemes / damping for the increment, the nonlinear model's for the
    trajectory (dyn_core_tlm.F90:740-926, sw_core_tlm.F90:1664-1682, 2341-2366)"""
    print(_run(True, 2, K=5, extra=TWO_SIDED))


@pytest.mark.gpu
def test_dyn_core_hydro_two_sided_gpu():
    _run(False, 2, K=5, extra=TWO_SIDED)


def test_dyn_core_hydro_beta_emu():
    """beta = 0.4: grad1_p_update (model/dyn_core_nlm.F90:1781-1872, TL dyn_core_tlm.F90:4163-4293) replaces one_grad_p"""
    print(_run(True, 3, extra=dict(beta=0.4)))


@pytest.mark.gpu
def test_dyn_core_hydro_beta_gpu():
    _run(False, 3, extra=dict(beta=0.4))


@pytest.mark.parametrize("beta", [0.0, 0.4])
def test_dyn_core_hydro_d_ext_emu(beta):
    """d_ext = 0.02 (the reference's default, model/fv_arrays_nlm.F90:328): external-mode divergence damping of the hydrostatic core
    (a2b_ord2 of delp, the delp-weighted column mean of d_sw's corner divergence, dyn_core_nlm.F90:642-726; the wk1 / wk2 terms of
    one_grad_p :1713-1771 and, with beta > 0, of grad1_p_update :1858-1867); the sponge layer runs nord = 0 (d_sw's own divergence)"""
    print(_run(True, 2, K=4, extra=dict(d_ext=0.02, beta=beta, n_sponge=2)))


@pytest.mark.gpu
def test_dyn_core_hydro_d_ext_gpu():
    _run(False, 2, K=4, extra=dict(d_ext=0.02, beta=0.4, n_sponge=2))


def test_dyn_core_hydro_d_ext_two_sided_emu():
    ts = dict(TWO_SIDED); ts["d_ext"] = 0.02
    print(_run(True, 2, K=5, extra=ts))


def test_dyn_core_hydro_heat_emu():
    """d_con = 1: heat source accumulated over the acoustic steps, filtered by del2_cubed, added to pt (dyn_core_nlm.F90:1052-1075)"""
    print(_run(True, 2, K=5, extra=dict(d_con=1.0)))     # layers 1-3 are sponge layers (d_con_k = 0): K = 5 leaves two heated ones


@pytest.mark.gpu
def test_del2_cubed_and_heat_gpu():
    _run_del2_cubed(False, 3)
    _run_heat_update(False, True); _run_heat_update(False, False)
    _run(False, 2, K=5, extra=dict(d_con=1.0))


@pytest.mark.gpu
@pytest.mark.parametrize("n_split", [1, 3])
def test_dyn_core_hydro_gpu(n_split):
    _run(False, n_split)


# ---- geopk on its own: the compute-domain form is compute_fv3_pressures (fvdyn.cu "compute_fv3_pressures") ----
def _run_geopk(emu, halo):
    from oracle.dyn_core import geopk
    N, K = 12, 6
    rng = np.random.default_rng(17)
    g = ograd(N)
    ptop, akap, cp_air = 100.0, 2.0 / 7.0, 1004.6
    f = {"delp": 1.0e4 / K * (1.0 + 0.2 * rng.random((6, K, N + 7, N + 7))), "pt": 300.0 + 5.0 * rnd(rng, N, K),
         "phis": 50.0 * rnd(rng, N, 1)}
    h = handle(N, K, emu)
    lo, hi = 1 - halo, N + halo
    outs = {"pk": (lo, hi, lo, hi), "gz": (lo, hi, lo, hi), "pe": (1, N, 1, N), "peln": (1, N, 1, N), "pkz": (1, N, 1, N)}
    hs = torch.from_numpy(f["phis"])
    return check_module(h, "geopk", N, K, f, ["delp", "pt"], outs, lambda dp, pt: geopk(dp, pt, hs, g, ptop, akap, cp_air, halo, False),
                        {"halo": halo, "cg": 0, "ptop": ptop, "akap": akap, "cp_air": cp_air}, rng,
                        out_nk={"pk": K + 1, "gz": K + 1, "pe": K + 1, "peln": K + 1, "pkz": K})


@pytest.mark.parametrize("halo", [0, 2])
def test_geopk_compute_fv3_pressures_emu(halo):
    print(_run_geopk(True, halo))


@pytest.mark.gpu
def test_geopk_compute_fv3_pressures_gpu():
    print(_run_geopk(False, 0))
