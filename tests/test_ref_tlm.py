"""Reference-derived pin of the ORACLE: the oracle's tangent (torch.func.jvp of its restated primal) against
statement-by-statement numpy transliterations of the reference's own Tapenade tangent routines (tests/ref_tlm/).  CPU only."""
import numpy as np
import pytest
import torch
from oracle import tp_core as otp
from common import metrics, ograd, rnd, relerr, region
from ref_tlm import F
from ref_tlm.xppm_tlm import xppm_tlm

TOL = 1e-13    # same arithmetic up to the order of a few additions


def _fa(a2d, N):
    """[NY, NX] array (Fortran (i, j) at [j + 2, i + 2]) -> F((isd, ied + 1), (jsd, jed + 1)) view indexed (i, j)"""
    return F((-2, N + 4), (-2, N + 4), data=np.ascontiguousarray(a2d.T))


@pytest.mark.parametrize("iord", [1, 2, 333])
def test_xppm_tlm_pins_oracle(iord):
    """XPPM_TLM (model_tlmadm/tp_core_tlm.F90:2328-2489) on whole cube tiles: is = 1, ie = N = npx - 1, both cube edges present"""
    N, K = 12, 2
    rng = np.random.default_rng(5)
    g = ograd(N)
    q = rnd(rng, N, K, 1.0, 10.0); c = rnd(rng, N, K, 0.4)
    dq = rnd(rng, N, K, 0.1); dc = rnd(rng, N, K, 0.01)
    j0, j1 = -2, N + 3                                   # jfirst:jlast = jsd:jed (the inner sweep of fv_tp_2d)
    fn = lambda a, b: otp.xppm(a, b, iord, g, j0, j1)
    flux_o, dflux_o = torch.func.jvp(fn, (torch.from_numpy(q), torch.from_numpy(c)), (torch.from_numpy(dq), torch.from_numpy(dc)))
    dxa = metrics(N)["dxa"]
    worst = 0.0
    for t in range(6):
        for k in range(K):
            fq, fq_tl, fc, fc_tl = _fa(q[t, k], N), _fa(dq[t, k], N), _fa(c[t, k], N), _fa(dc[t, k], N)
            # q(isd:ied, jfirst:jlast), c(is:ie+1, ...): the wider F arrays hold them at the same Fortran indices
            flux, flux_tl = xppm_tlm(fq, fq_tl, fc, fc_tl, iord, 1, N, -2, N + 3, j0, j1, -2, N + 3, N + 1, N + 1, _fa(dxa[t], N))
            ref = flux.a.T          # [j - j0, i - 1]
            ref_tl = flux_tl.a.T
            o = region(flux_o[t, k].numpy(), 1, N + 1, j0, j1)
            o_tl = region(dflux_o[t, k].numpy(), 1, N + 1, j0, j1)
            worst = max(worst, relerr(o, ref), relerr(o_tl, ref_tl))
            assert np.abs(ref_tl).max() > 0
    assert worst <= TOL, worst


def test_sim1_solver_tlm_pins_oracle():
    """SIM1_SOLVER_TLM (model_tlmadm/nh_utils_tlm.F90:2548-2762) on 40 columns of 72 layers with atmosphere-like magnitudes"""
    from oracle import nh as onh
    from ref_tlm.sim1_solver_tlm import sim1_solver_tlm
    km, ni = 72, 40
    rng = np.random.default_rng(11)
    rgas, kappa = 287.05, 287.05 / 1004.6
    gama = 1. / (1. - kappa)
    dt, p_fac = 25.0, 0.05
    dp = np.linspace(2.0, 1600.0, km)[:, None] * (1.0 + 0.05 * rng.standard_normal((km, ni)))      # Pa
    dm2 = dp / 9.80665
    pem = np.concatenate([np.full((1, ni), 1.0), 1.0 + np.cumsum(dp, axis=0)], axis=0)
    pm2 = (pem[1:] - pem[:-1]) / np.log(pem[1:] / pem[:-1])
    pt2 = (300.0 + 5.0 * rng.standard_normal((km, ni)))
    dz2 = -(dm2 * rgas * pt2 / pm2 ** (1. - kappa) * pm2 ** (-kappa) * 1.0) * (pm2 ** (kappa) / pm2 ** kappa)   # ~ -dm R T / p (hydrostatic)
    dz2 = dz2 * (1.0 + 0.01 * rng.standard_normal((km, ni)))
    w2 = 0.5 * rng.standard_normal((km, ni))
    ws = 0.1 * rng.standard_normal(ni)
    pert = lambda a, s: s * np.abs(a).mean() * rng.standard_normal(a.shape)
    d = dict(dm2=pert(dm2, 1e-3), pm2=pert(pm2, 1e-3), pem=pert(pem, 1e-3), w2=pert(w2, 1e-1), dz2=pert(dz2, 1e-3), pt2=pert(pt2, 1e-3), ws=pert(ws, 1e-1))
    pe, pe_tl, w2n, w2n_tl, dzn, dzn_tl = sim1_solver_tlm(dt, km, rgas, gama, kappa, dm2, d["dm2"], pm2, d["pm2"], pem, d["pem"], w2, d["w2"],
                                                          dz2, d["dz2"], pt2, d["pt2"], ws, d["ws"], p_fac)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))

    def fn(dm2_, pm2_, pem_, w2_, dz2_, pt2_, ws_):
        w, p, z = onh.sim1_solver(dt, list(dm2_), list(pm2_), list(pem_), list(w2_), list(dz2_), list(pt2_), ws_, rgas, gama, kappa, p_fac)
        return torch.stack(w), torch.stack(p), torch.stack(z)
    prim = (T(dm2), T(pm2), T(pem), T(w2), T(dz2), T(pt2), T(ws))
    tang = (T(d["dm2"]), T(d["pm2"]), T(d["pem"]), T(d["w2"]), T(d["dz2"]), T(d["pt2"]), T(d["ws"]))
    (w_o, p_o, z_o), (dw_o, dp_o, dz_o) = torch.func.jvp(fn, prim, tang)
    errs = dict(w2=relerr(w_o.numpy(), w2n), pe=relerr(p_o.numpy(), pe), dz2=relerr(z_o.numpy(), dzn),
                w2_tl=relerr(dw_o.numpy(), w2n_tl), pe_tl=relerr(dp_o.numpy(), pe_tl), dz2_tl=relerr(dz_o.numpy(), dzn_tl))
    assert np.abs(w2n_tl).max() > 0 and np.abs(dzn_tl).max() > 0
    print('sim1 errs', errs)
    assert max(errs.values()) <= 1e-12, errs      # achieved 1e-13 (exp / log round-off through 72 levels)


def test_a2b_ord4_tlm_pins_oracle():
    """A2B_ORD4_TLM (model_tlmadm/a2b_edge_tlm.F90:546-1158, with EXTRAP_CORNER_TLM :1491) on whole cube tiles: all four corners, all four
    edges, the interior; the operator is linear, so the oracle's jvp must be the operator applied to the perturbation"""
    from oracle.a2b_edge import a2b_ord4
    from ref_tlm.a2b_ord4_tlm import a2b_ord4_tlm
    N, K = 12, 1
    rng = np.random.default_rng(23)
    g = ograd(N)
    M = metrics(N)
    q = rnd(rng, N, K, 1.0, 5.0); dq = rnd(rng, N, K, 0.3)
    fn = lambda a: a2b_ord4(a, g)
    out_o, dout_o = torch.func.jvp(fn, (torch.from_numpy(q),), (torch.from_numpy(dq),))
    npx = npy = N + 1
    worst = 0.0
    for t in range(6):
        P = lambda a, i, j: a[t, j + 2, i + 2]                # Fortran (i, j) of a [6, NY, NX(, 2)] metric
        qout, qout_tl = a2b_ord4_tlm(_fa(q[t, 0], N), _fa(dq[t, 0], N), lambda i, j: P(M["grid"], i, j), lambda i, j: P(M["agrid"], i, j),
                                     _fa(M["dxa"][t], N), _fa(M["dya"][t], N),
                                     lambda j: M["edge_w"][t, j + 2], lambda j: M["edge_e"][t, j + 2], lambda i: M["edge_s"][t, i + 2], lambda i: M["edge_n"][t, i + 2],
                                     npx, npy, 1, N, 1, N, 3)
        ref = qout.a.T; ref_tl = qout_tl.a.T                   # [j + 2, i + 2]
        o = region(out_o[t, 0].numpy(), 1, N + 1, 1, N + 1); o_tl = region(dout_o[t, 0].numpy(), 1, N + 1, 1, N + 1)
        worst = max(worst, relerr(o, region(ref, 1, N + 1, 1, N + 1)), relerr(o_tl, region(ref_tl, 1, N + 1, 1, N + 1)))
        assert np.abs(region(ref_tl, 1, N + 1, 1, N + 1)).min() > 0        # every corner point of the tile is defined
    print("a2b_ord4 worst", worst)
    assert worst <= TOL, worst


@pytest.mark.parametrize("case", ["plain_ord2", "plain_ord333", "plain_ord1", "damp_nord2", "damp_nord0", "mass_flux", "mass_flux_damp_nord1"])
def test_fv_tp_2d_tlm_pins_oracle(case):
    """FV_TP_2D_TLM with YPPM_TLM, COPY_CORNERS_TLM and DELN_FLUX_TLM (model_tlmadm/tp_core_tlm.F90:2123-2924) on two whole cube tiles:
    the inner / outer sweeps, the corner copies between them, the flux averaging with either mass-flux convention and both del-n
    branches (damp * q without mass, mass-weighted with it)."""
    from ref_tlm.fv_tp_2d_tlm import fv_tp_2d_tlm, BD
    N, K = 12, 1
    rng = np.random.default_rng(23)
    M = metrics(N); g = ograd(N)
    hord = dict(plain_ord333=333, plain_ord1=1).get(case, 2)
    mf = case.startswith("mass_flux")
    nord, damp_c = dict(damp_nord2=(2, 0.06), damp_nord0=(0, 0.12), mass_flux_damp_nord1=(1, 0.05)).get(case, (None, None))
    area = M["area"][:, None]
    names = ["q", "crx", "cry", "xfx", "yfx", "ra_x", "ra_y"] + (["mfx", "mfy", "mass"] if mf else [])
    v = dict(q=rnd(rng, N, K, 1.0, 10.0), crx=rnd(rng, N, K, 0.4), cry=rnd(rng, N, K, 0.4),
             xfx=area * rnd(rng, N, K, 0.3), yfx=area * rnd(rng, N, K, 0.3),
             ra_x=area * (1.0 + 0.1 * rnd(rng, N, K)), ra_y=area * (1.0 + 0.1 * rnd(rng, N, K)),
             mfx=area * rnd(rng, N, K, 30.0), mfy=area * rnd(rng, N, K, 30.0), mass=rnd(rng, N, K, 5.0, 100.0))
    d = {n: 1e-2 * np.abs(v[n]).mean() * rnd(rng, N, K) for n in names}

    def fn(*a):
        kw = dict(zip(names, a))
        fx, fy, _ = otp.fv_tp_2d_damp(kw["q"], kw["crx"], kw["cry"], hord, kw["xfx"], kw["yfx"], g, kw["ra_x"], kw["ra_y"],
                                      kw.get("mfx"), kw.get("mfy"), kw.get("mass"),
                                      None if nord is None else [nord] * K, None if nord is None else [damp_c] * K)
        return fx, fy
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    (fx_o, fy_o), (dfx_o, dfy_o) = torch.func.jvp(fn, tuple(T(v[n]) for n in names), tuple(T(d[n]) for n in names))
    bd = BD(N)
    worst = {}
    for t in (0, 4):
        gs = {k: _fa(M[k][t], N) for k in ("area", "rarea", "dxa", "dya", "del6_u", "del6_v")}
        gs["da_min"] = M["da_min"]
        a = {n: _fa(v[n][t, 0], N) for n in names}
        a_tl = {n: _fa(d[n][t, 0], N) for n in names}
        opt = {}
        if mf:
            opt = dict(mfx=a["mfx"], mfx_tl=a_tl["mfx"], mfy=a["mfy"], mfy_tl=a_tl["mfy"])
            if nord is not None:
                opt.update(mass=a["mass"], mass_tl=a_tl["mass"])
        if nord is not None:
            opt.update(nord=nord, damp_c=damp_c)
        fx, fx_tl, fy, fy_tl = fv_tp_2d_tlm(a["q"], a_tl["q"], a["crx"], a_tl["crx"], a["cry"], a_tl["cry"], N + 1, N + 1, hord,
                                            a["xfx"], a_tl["xfx"], a["yfx"], a_tl["yfx"], gs, bd, a["ra_x"], a_tl["ra_x"],
                                            a["ra_y"], a_tl["ra_y"], **opt)
        for nm, ref, o, rg in (("fx", fx, fx_o, (1, N + 1, 1, N)), ("fx_tl", fx_tl, dfx_o, (1, N + 1, 1, N)),
                               ("fy", fy, fy_o, (1, N, 1, N + 1)), ("fy_tl", fy_tl, dfy_o, (1, N, 1, N + 1))):
            assert np.abs(ref.a).max() > 0
            worst[nm] = max(worst.get(nm, 0.0), relerr(region(o[t, 0].numpy(), *rg), ref.a.T))
    print("fv_tp_2d", case, worst)
    assert max(worst.values()) <= TOL, worst


@pytest.mark.parametrize("iord", [1, 2, 333])
def test_xtp_u_ytp_v_tlm_pin_oracle(iord):
    """XTP_U_TLM / YTP_V_TLM (model_tlmadm/sw_core_tlm.F90:7272-7757) on two whole cube tiles: the momentum transport of d_sw with its
    cube-edge one-sided values and the zeroed corner rows (j = 1, npy for u; i = 1, npx for v)."""
    from oracle import d_sw as odsw
    from ref_tlm.xtp_ytp_tlm import xtp_u_tlm, ytp_v_tlm
    N, K = 12, 1
    rng = np.random.default_rng(31)
    M = metrics(N); g = ograd(N)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    worst = {}
    for nm, fn_o, fn_r, dn, rdn in (("xtp_u", odsw.xtp_u, xtp_u_tlm, "dx", "rdx"), ("ytp_v", odsw.ytp_v, ytp_v_tlm, "dy", "rdy")):
        w = rnd(rng, N, K, 5.0, 10.0); c = M[dn][:, None] * rnd(rng, N, K, 0.4)
        dw = rnd(rng, N, K, 0.1); dc = M[dn][:, None] * rnd(rng, N, K, 0.01)
        f_o, df_o = torch.func.jvp(lambda c_, w_: fn_o(c_, w_, g, iord), (T(c), T(w)), (T(dc), T(dw)))
        for t in (1, 5):
            flux, flux_tl = fn_r(1, N, 1, N, -2, N + 3, -2, N + 3, _fa(c[t, 0], N), _fa(dc[t, 0], N), _fa(w[t, 0], N), _fa(dw[t, 0], N), iord,
                                 _fa(M[dn][t], N), _fa(M[rdn][t], N), N + 1, N + 1)
            assert np.abs(flux_tl.a).max() > 0
            worst[nm] = max(worst.get(nm, 0.0), relerr(region(f_o[t, 0].numpy(), 1, N + 1, 1, N + 1), flux.a.T))
            worst[nm + "_tl"] = max(worst.get(nm + "_tl", 0.0), relerr(region(df_o[t, 0].numpy(), 1, N + 1, 1, N + 1), flux_tl.a.T))
    print("xtp/ytp", iord, worst)
    assert max(worst.values()) <= TOL, worst


def _fa3(a3d, N):
    """[K, NY, NX] -> F((isd, ied + 1), (jsd, jed + 1), (1, K)) indexed (i, j, k)"""
    return F((-2, N + 4), (-2, N + 4), (1, a3d.shape[0]), data=np.ascontiguousarray(a3d.transpose(2, 1, 0)))


def test_update_dz_d_tlm_pins_oracle():
    """UPDATE_DZ_D_TLM (model_tlmadm/nh_utils_tlm.F90:381-588) with EDGE_PROFILE_TLM (:3319-3472) and DEL6_VT_FLUX_TLM
    (sw_core_tlm.F90:3621-3728) on one whole cube tile: interface Courant numbers / fluxes, transport of the interface heights with and
    without del-n damping, the surface vertical velocity and the minimum-thickness limiter (a thin layer makes it active)."""
    from oracle import nh as onh
    from ref_tlm.update_dz_tlm import update_dz_d_tlm
    from ref_tlm.fv_tp_2d_tlm import BD
    N, K = 12, 5
    rng = np.random.default_rng(41)
    M = metrics(N); g = ograd(N)
    area = M["area"][:, None]
    dp0 = np.array([300., 700., 1500., 2500., 3000.])
    ndif = [2, 1, 0, 0, 0]; dampc = [0.06, 0.05, 0.04, 0.0, 0.0]
    damp = [(d * M["da_min"]) ** (n + 1) if d > 0 else 0.0 for n, d in zip(ndif, dampc)]
    hord, rdt = 2, 1. / 30.
    NX = N + 7
    zs = 100.0 * rng.standard_normal((6, 1, NX, NX))
    thick = np.array([900., 800., 1.0, 600., 500., 0.0])[::-1].cumsum()[::-1]          # layer 3 is 1 m thick: limiter active there
    zh = zs + thick.reshape(1, K + 1, 1, 1) + 1.5 * rng.standard_normal((6, K + 1, NX, NX))
    v = dict(zh=zh, crx=rnd(rng, N, K, 0.3), cry=rnd(rng, N, K, 0.3), xfx=area * rnd(rng, N, K, 0.2), yfx=area * rnd(rng, N, K, 0.2))
    names = list(v)
    d = {n: 1e-2 * max(np.abs(v[n]).std(), 1e-30) * rng.standard_normal(v[n].shape) for n in names}
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))

    def fn(zh_, crx_, cry_, xfx_, yfx_):
        return onh.update_dz_d(ndif, damp, hord, dp0, T(zs), zh_, crx_, cry_, xfx_, yfx_, g, rdt)
    (zh_o, ws_o), (dzh_o, dws_o) = torch.func.jvp(fn, tuple(T(v[n]) for n in names), tuple(T(d[n]) for n in names))
    t = 2
    gs = {k: _fa(M[k][t], N) for k in ("area", "rarea", "dxa", "dya", "del6_u", "del6_v")}
    gs["da_min"] = M["da_min"]
    a = {n: _fa3(v[n][t], N) for n in names}; a_tl = {n: _fa3(d[n][t], N) for n in names}
    ws, ws_tl = update_dz_d_tlm(ndif + [0], damp + [0.0], hord, 1, N, 1, N, K, 3, N + 1, N + 1, F((1, K), data=dp0), _fa(zs[t, 0], N),
                                a["zh"], a_tl["zh"], a["crx"], a_tl["crx"], a["cry"], a_tl["cry"], a["xfx"], a_tl["xfx"],
                                a["yfx"], a_tl["yfx"], rdt, gs, BD(N))
    own = lambda x: region(x, 1, N, 1, N)
    ref_zh = a["zh"].a.transpose(2, 1, 0); ref_zh_tl = a_tl["zh"].a.transpose(2, 1, 0)
    limited = (own(ref_zh)[:-1] - own(ref_zh)[1:] == 2.0).mean()
    assert 0.05 < limited < 0.5, limited                # the dz_min branch is exercised, and not everywhere
    errs = dict(zh=relerr(own(zh_o[t].numpy()), own(ref_zh)), zh_tl=relerr(own(dzh_o[t].numpy()), own(ref_zh_tl)),
                ws=relerr(own(ws_o[t, 0].numpy()), ws.a.T), ws_tl=relerr(own(dws_o[t, 0].numpy()), ws_tl.a.T))
    print("update_dz_d", errs, "limited", limited)
    assert np.abs(ws_tl.a).max() > 0
    assert max(errs.values()) <= TOL, errs


def test_update_dz_c_tlm_pins_oracle():
    """UPDATE_DZ_C_TLM (model_tlmadm/nh_utils_tlm.F90:51-233) with FILL_4CORNERS_TLM (sw_core_tlm.F90:7138-7211) on two whole cube tiles:
    interface fluxes from the layer winds, first-order upwind transport of the heights over (0:npx, 0:npy) and the dz_min limiter."""
    from oracle import nh as onh
    from ref_tlm.update_dz_tlm import update_dz_c_tlm
    N, K = 12, 5
    rng = np.random.default_rng(43)
    M = metrics(N); g = ograd(N)
    area = M["area"][:, None]
    dp0 = np.array([300., 700., 1500., 2500., 3000.])
    dt = 15.0
    NX = N + 7
    zs = 100.0 * rng.standard_normal((6, 1, NX, NX))
    thick = np.array([900., 800., 1.0, 600., 500., 0.0])[::-1].cumsum()[::-1]
    gz = zs + thick.reshape(1, K + 1, 1, 1) + 1.5 * rng.standard_normal((6, K + 1, NX, NX))
    v = dict(ut=area * rnd(rng, N, K, 0.2), vt=area * rnd(rng, N, K, 0.2), gz=gz)
    names = list(v)
    d = {n: 1e-2 * np.abs(v[n]).std() * rng.standard_normal(v[n].shape) for n in names}
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    (gz_o, ws_o), (dgz_o, dws_o) = torch.func.jvp(lambda ut_, vt_, gz_: onh.update_dz_c(dt, dp0, T(zs), ut_, vt_, gz_, g),
                                                  tuple(T(v[n]) for n in names), tuple(T(d[n]) for n in names))
    errs = {}
    for t in (0, 3):
        a = {n: _fa3(v[n][t], N) for n in names}; a_tl = {n: _fa3(d[n][t], N) for n in names}
        ws, ws_tl = update_dz_c_tlm(1, N, 1, N, K, 3, dt, F((1, K), data=dp0), _fa(zs[t, 0], N), _fa(M["area"][t], N),
                                    a["ut"], a_tl["ut"], a["vt"], a_tl["vt"], a["gz"], a_tl["gz"], N + 1, N + 1)
        ext = lambda x: region(x, 0, N + 1, 0, N + 1)
        ref = dict(gz=a["gz"].a.transpose(2, 1, 0), gz_tl=a_tl["gz"].a.transpose(2, 1, 0),
                   ws=np.pad(ws.a.T, ((0, 1), (0, 1))), ws_tl=np.pad(ws_tl.a.T, ((0, 1), (0, 1))))
        got = dict(gz=gz_o[t].numpy(), gz_tl=dgz_o[t].numpy(), ws=ws_o[t, 0].numpy(), ws_tl=dws_o[t, 0].numpy())
        for k in ref:
            errs[k] = max(errs.get(k, 0.0), relerr(ext(got[k]), ext(ref[k])))
        assert np.abs(ws_tl.a).max() > 0
    print("update_dz_c", errs)
    assert max(errs.values()) <= TOL, errs


def test_riem_solver_tlm_pins_oracle():
    """RIEM_SOLVER_C_TLM (model_tlmadm/nh_utils_tlm.F90:723-845) and RIEM_SOLVER3_TLM (model_tlmadm/nh_core_tlm.F90:49-243), SIM1 branch,
    on the columns of one cube tile with atmosphere-like profiles (10 layers)."""
    from oracle import nh as onh
    from test_nh import column_state, NHCFG
    from ref_tlm.riem_solver_tlm import riem_solver_c_tlm, riem_solver3_tlm
    N, K = 12, 10
    f, rng, ak, bk = column_state(N, K, 17)
    cfg = dict(NHCFG); cfg["a_imp"] = 1.0
    dts = 75.0
    names = ["delp", "pt", "z", "w", "ws"]
    d = {n: 1e-3 * np.abs(f[n]).std() * rng.standard_normal(f[n].shape) for n in names}
    d["w"] *= 100.0; d["ws"] *= 100.0
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    t = 0
    col = lambda a: np.ascontiguousarray(a[t].reshape(a.shape[1], -1))        # [K(+1), ncol]
    ws_c = lambda a: col(a)[0]
    zb = T(f["zb"])
    # C-grid solver: heights in, geopotential out, hs = grav * zb
    (pef_o, gz_o), (dpef_o, dgz_o) = torch.func.jvp(
        lambda delp, pt, z, w, ws: onh.riem_solver_c(dts, delp, pt, z, w, ws, zb * cfg["grav"], cfg),
        tuple(T(f[n]) for n in names), tuple(T(d[n]) for n in names))
    gz, gz_tl, pef, pef_tl = riem_solver_c_tlm(cfg["grav"], cfg["rdgas"], dts, K, cfg["akap"], cfg["ptop"], ws_c(f["zb"]) * cfg["grav"],
                                               col(f["w"]), col(d["w"]), col(f["pt"]), col(d["pt"]), col(f["delp"]), col(d["delp"]),
                                               col(f["z"]), col(d["z"]), ws_c(f["ws"]), ws_c(d["ws"]), cfg["p_fac"])
    errs = dict(c_gz=relerr(col(gz_o.numpy()), gz), c_gz_tl=relerr(col(dgz_o.numpy()), gz_tl),
                c_pef=relerr(col(pef_o.numpy()), pef), c_pef_tl=relerr(col(dpef_o.numpy()), pef_tl))
    # D-grid solver
    (w_o, dz_o, zh_o, pp_o), (dw_o, ddz_o, dzh_o, dpp_o) = torch.func.jvp(
        lambda delp, pt, z, w, ws: onh.riem_solver3(dts, delp, pt, z, w, ws, zb, cfg),
        tuple(T(f[n]) for n in names), tuple(T(d[n]) for n in names))
    r = riem_solver3_tlm(cfg["grav"], cfg["rdgas"], dts, K, cfg["akap"], cfg["ptop"], ws_c(f["zb"]), col(f["w"]), col(d["w"]),
                         col(f["pt"]), col(d["pt"]), col(f["delp"]), col(d["delp"]), col(f["z"]), col(d["z"]), ws_c(f["ws"]), ws_c(d["ws"]),
                         cfg["p_fac"])
    for nm, o, do in (("w", w_o, dw_o), ("delz", dz_o, ddz_o), ("zh", zh_o, dzh_o), ("ppe", pp_o, dpp_o)):
        errs["d_" + nm] = relerr(col(o.numpy()), r[nm]); errs["d_" + nm + "_tl"] = relerr(col(do.numpy()), r[nm + "_tl"])
        assert np.abs(r[nm + "_tl"]).max() > 0
    print("riem", errs)
    assert max(errs.values()) <= TOL, errs          # achieved 7e-15


def test_nh_p_grad_tlm_pins_oracle():
    """NH_P_GRAD_TLM (model_tlmadm/dyn_core_tlm.F90:3340-3470) on one whole cube tile: a2b_ord4 of pp / pk / gz / delp, the hydrostatic and
    the non-hydrostatic pressure-gradient increments of the D-grid winds."""
    from oracle import dyn_core as odc
    from ref_tlm.nh_p_grad_tlm import nh_p_grad_tlm
    from ref_tlm.a2b_ord4_tlm import a2b_ord4_tlm
    from ref_tlm.fv_tp_2d_tlm import BD
    N, K = 12, 3
    rng = np.random.default_rng(53)
    M = metrics(N); g = ograd(N)
    NX = N + 7
    akap, dt, ptk = 2. / 7., 20.0, 100.0 ** (2. / 7.)
    pe = np.array([100., 3.0e4, 7.0e4, 1.0e5]).reshape(1, K + 1, 1, 1) * (1.0 + 0.01 * rng.standard_normal((6, K + 1, NX, NX)))
    v = dict(u=rnd(rng, N, K, 10.0) * M["dx"][:, None], v=rnd(rng, N, K, 10.0) * M["dy"][:, None], pk=pe ** akap,
             gz=9.80665 * (np.array([30000., 9000., 3000., 0.]).reshape(1, K + 1, 1, 1) + 50.0 * rng.standard_normal((6, K + 1, NX, NX))),
             pp=30.0 * rng.standard_normal((6, K + 1, NX, NX)), delp=pe[:, 1:] - pe[:, :-1])
    names = list(v)
    d = {n: 1e-2 * np.abs(v[n]).std() * rng.standard_normal(v[n].shape) for n in names}
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    (u_o, v_o), (du_o, dv_o) = torch.func.jvp(lambda u, v_, pk, gz, pp, delp: odc.grad_p(u, v_, pk, gz, g, dt, ptk, pp=pp, delp=delp),
                                              tuple(T(v[n]) for n in names), tuple(T(d[n]) for n in names))
    t = 4
    P = lambda a, i, j: a[t, j + 2, i + 2]
    dxa, dya = F((-2, N + 3), (-2, N + 3), data=M["dxa"][t].T[:-1, :-1]), F((-2, N + 3), (-2, N + 3), data=M["dya"][t].T[:-1, :-1])
    a2b = lambda qin, qin_tl: a2b_ord4_tlm(qin, qin_tl, lambda i, j: P(M["grid"], i, j), lambda i, j: P(M["agrid"], i, j), dxa, dya,
                                           lambda j: M["edge_w"][t, j + 2], lambda j: M["edge_e"][t, j + 2],
                                           lambda i: M["edge_s"][t, i + 2], lambda i: M["edge_n"][t, i + 2], N + 1, N + 1, 1, N, 1, N, 3)
    a = {n: _fa3(v[n][t], N) for n in names}; a_tl = {n: _fa3(d[n][t], N) for n in names}
    nh_p_grad_tlm(a["u"], a_tl["u"], a["v"], a_tl["v"], a["pp"], a_tl["pp"], a["gz"], a_tl["gz"], a["delp"], a_tl["delp"], a["pk"], a_tl["pk"],
                  dt, _fa(M["rdx"][t], N), _fa(M["rdy"][t], N), BD(N), K, ptk, a2b)
    back = lambda f: f.a.transpose(2, 1, 0)
    errs = dict(u=relerr(region(u_o[t].numpy(), 1, N, 1, N + 1), region(back(a["u"]), 1, N, 1, N + 1)),
                u_tl=relerr(region(du_o[t].numpy(), 1, N, 1, N + 1), region(back(a_tl["u"]), 1, N, 1, N + 1)),
                v=relerr(region(v_o[t].numpy(), 1, N + 1, 1, N), region(back(a["v"]), 1, N + 1, 1, N)),
                v_tl=relerr(region(dv_o[t].numpy(), 1, N + 1, 1, N), region(back(a_tl["v"]), 1, N + 1, 1, N)))
    print("nh_p_grad", errs)
    assert max(errs.values()) <= TOL, errs


@pytest.mark.parametrize("hydrostatic", [False, True])
def test_p_grad_c_tlm_pins_oracle(hydrostatic):
    """P_GRAD_C_TLM (model_tlmadm/dyn_core_tlm.F90:3194-3275) on one whole cube tile, both weightings (delpc / pkc differences)"""
    from oracle import dyn_core as odc
    from ref_tlm.p_grad_c_tlm import p_grad_c_tlm
    from ref_tlm.fv_tp_2d_tlm import BD
    N, K = 12, 3
    rng = np.random.default_rng(59)
    M = metrics(N); g = ograd(N)
    NX = N + 7
    dt2 = 10.0
    pe = np.array([100., 3.0e4, 7.0e4, 1.0e5]).reshape(1, K + 1, 1, 1) * (1.0 + 0.01 * rng.standard_normal((6, K + 1, NX, NX)))
    v = dict(delpc=pe[:, 1:] - pe[:, :-1], pkc=(pe ** (2. / 7.) if hydrostatic else pe),
             gz=9.80665 * (np.array([30000., 9000., 3000., 0.]).reshape(1, K + 1, 1, 1) + 50.0 * rng.standard_normal((6, K + 1, NX, NX))),
             uc=rnd(rng, N, K, 10.0), vc=rnd(rng, N, K, 10.0))
    names = list(v)
    d = {n: 1e-2 * np.abs(v[n]).std() * rng.standard_normal(v[n].shape) for n in names}
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    (u_o, v_o), (du_o, dv_o) = torch.func.jvp(lambda delpc, pkc, gz, uc, vc: odc.p_grad_c(dt2, delpc, pkc, gz, uc, vc, g, hydrostatic),
                                              tuple(T(v[n]) for n in names), tuple(T(d[n]) for n in names))
    t = 3
    a = {n: _fa3(v[n][t], N) for n in names}; a_tl = {n: _fa3(d[n][t], N) for n in names}
    p_grad_c_tlm(dt2, K, a["delpc"], a_tl["delpc"], a["pkc"], a_tl["pkc"], a["gz"], a_tl["gz"], a["uc"], a_tl["uc"], a["vc"], a_tl["vc"],
                 BD(N), _fa(M["rdxc"][t], N), _fa(M["rdyc"][t], N), hydrostatic)
    back = lambda f: f.a.transpose(2, 1, 0)
    errs = dict(uc=relerr(region(u_o[t].numpy(), 1, N + 1, 1, N), region(back(a["uc"]), 1, N + 1, 1, N)),
                uc_tl=relerr(region(du_o[t].numpy(), 1, N + 1, 1, N), region(back(a_tl["uc"]), 1, N + 1, 1, N)),
                vc=relerr(region(v_o[t].numpy(), 1, N, 1, N + 1), region(back(a["vc"]), 1, N, 1, N + 1)),
                vc_tl=relerr(region(dv_o[t].numpy(), 1, N, 1, N + 1), region(back(a_tl["vc"]), 1, N, 1, N + 1)))
    print("p_grad_c", hydrostatic, errs)
    assert max(errs.values()) <= TOL, errs


def test_d2a2c_vect_divergence_corner_tlm_pin_oracle():
    """D2A2C_VECT_TLM (model_tlmadm/sw_core_tlm.F90:5874-6395) and DIVERGENCE_CORNER_TLM (:3805-3966) on two whole cube tiles: the D -> A -> C
    wind interpolation with its edge / corner special cases and upstream sin_sg choice, then the corner divergence from its outputs."""
    from oracle import sw_core as osw
    from ref_tlm.d2a2c_vect_tlm import d2a2c_vect_tlm, divergence_corner_tlm
    from ref_tlm.fv_tp_2d_tlm import BD
    N, K = 12, 1
    rng = np.random.default_rng(61)
    M = metrics(N); g = ograd(N)
    u = rnd(rng, N, K, 10.0, 3.0); v = rnd(rng, N, K, 10.0, -2.0)
    du = rnd(rng, N, K, 0.1); dv = rnd(rng, N, K, 0.1)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))

    def fn(u_, v_):
        ua, va, uc, vc, ut, vt = osw.d2a2c_vect(u_, v_, g, True)
        return ua, va, uc, vc, ut, vt, osw.divergence_corner(u_, v_, ua, va, g)
    out_o, dout_o = torch.func.jvp(fn, (T(u), T(v)), (T(du), T(dv)))
    onames = ["ua", "va", "uc", "vc", "ut", "vt", "divg_d"]
    # regions the reference defines (and the callers read): ua / va with dord4 on is-2..ie+2, uc / ut on 0..npx+1 x 0..npy-1+1, ...
    regs = dict(ua=(-1, N + 2, -1, N + 2), va=(-1, N + 2, -1, N + 2), uc=(0, N + 2, 0, N + 1), ut=(0, N + 2, 0, N + 1),
                vc=(0, N + 1, 0, N + 2), vt=(0, N + 1, 0, N + 2), divg_d=(1, N + 1, 1, N + 1))
    errs = {}
    for t in (0, 5):
        gs = {k: _fa(M[k][t], N) for k in ("cosa_u", "cosa_v", "cosa_s", "rsin_u", "rsin_v", "rsin2", "dxa", "dya", "dxc", "dyc", "rarea_c")}
        gs["sin_sg"] = lambda i, j, n, t=t: M["sin_sg"][t, j + 2, i + 2, n]
        gs["cos_sg"] = lambda i, j, n, t=t: M["cos_sg"][t, j + 2, i + 2, n]
        fu, fu_tl, fv, fv_tl = _fa(u[t, 0], N), _fa(du[t, 0], N), _fa(v[t, 0], N), _fa(dv[t, 0], N)
        r = d2a2c_vect_tlm(fu, fu_tl, fv, fv_tl, True, gs, BD(N), N + 1, N + 1)
        r["divg_d"], r["divg_d_tl"] = divergence_corner_tlm(fu, fu_tl, fv, fv_tl, r["ua"], r["ua_tl"], r["va"], r["va_tl"], gs, BD(N), N + 1, N + 1)
        for k, nm in enumerate(onames):
            for sfx, src in (("", out_o), ("_tl", dout_o)):
                ref = r[nm + sfx].a.T
                ref = np.pad(ref, ((0, N + 7 - ref.shape[0]), (0, N + 7 - ref.shape[1])))
                errs[nm + sfx] = max(errs.get(nm + sfx, 0.0), relerr(region(src[k][t, 0].numpy(), *regs[nm]), region(ref, *regs[nm])))
                assert np.abs(region(ref, *regs[nm])).max() > 0 and np.abs(region(ref, *regs[nm])).max() < 1e20
    print("d2a2c / divergence_corner", errs)
    assert max(errs.values()) <= TOL, errs


@pytest.mark.parametrize("hydrostatic", [False, True])
def test_c_sw_tlm_pins_oracle(hydrostatic):
    """C_SW_TLM (model_tlmadm/sw_core_tlm.F90:87-645) on two whole cube tiles: the complete C-grid half step (winds on A / C grids, corner
    divergence, upwind transport of delp / pt / w, kinetic energy, absolute vorticity and the time-centred C-grid winds)."""
    from oracle import sw_core as osw
    from test_c_sw import smooth_state
    from ref_tlm.c_sw_tlm import c_sw_tlm
    from ref_tlm.fv_tp_2d_tlm import BD
    N, K = 12, 1
    f, rng = smooth_state(N, K, 71)
    M = metrics(N); g = ograd(N)
    dt2 = 225.0
    names = ["delp", "pt", "u", "v", "w"]
    d = {n: 1e-2 * np.abs(f[n]).std() * rnd(rng, N, K) for n in names}
    onames = ["delpc", "ptc", "uc", "vc", "ua", "va", "ut", "vt", "divg_d"] + ([] if hydrostatic else ["wc"])
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))

    def fn(*a):
        o = osw.c_sw(*a, g, dt2, hydrostatic, 1)
        return tuple(o[k] for k in onames)
    out_o, dout_o = torch.func.jvp(fn, tuple(T(f[n]) for n in names), tuple(T(d[n]) for n in names))
    npx = N + 1
    regs = {"delpc": (0, npx, 0, npx), "ptc": (0, npx, 0, npx), "wc": (0, npx, 0, npx), "uc": (1, npx, 1, N), "vc": (1, N, 1, npx),
            "ua": (0, npx, 0, npx), "va": (0, npx, 0, npx), "ut": (0, npx + 1, 0, npx), "vt": (0, npx, 0, npx + 1), "divg_d": (1, npx, 1, npx)}
    errs = {}
    for t in (1, 4):
        gs = {k: _fa(M[k][t], N) for k in ("cosa_u", "cosa_v", "cosa_s", "sina_u", "sina_v", "rsin_u", "rsin_v", "rsin2", "dxa", "dya", "dx", "dy",
                                           "dxc", "dyc", "rarea", "rarea_c", "rdxc", "rdyc", "fC")}
        gs["sin_sg"] = lambda i, j, n, t=t: M["sin_sg"][t, j + 2, i + 2, n]
        gs["cos_sg"] = lambda i, j, n, t=t: M["cos_sg"][t, j + 2, i + 2, n]
        a = {n: _fa(f[n][t, 0], N) for n in names}; a_tl = {n: _fa(d[n][t, 0], N) for n in names}
        r = c_sw_tlm(a["delp"], a_tl["delp"], a["pt"], a_tl["pt"], a["u"], a_tl["u"], a["v"], a_tl["v"], a["w"], a_tl["w"], 1, dt2, hydrostatic,
                     True, BD(N), gs, npx, npx)
        for k, nm in enumerate(onames):
            for sfx, src in (("", out_o), ("_tl", dout_o)):
                ref = r[nm + sfx].a.T
                ref = np.pad(ref, ((0, N + 7 - ref.shape[0]), (0, N + 7 - ref.shape[1])))
                errs[nm + sfx] = max(errs.get(nm + sfx, 0.0), relerr(region(src[k][t, 0].numpy(), *regs[nm]), region(ref, *regs[nm])))
                assert np.abs(region(ref, *regs[nm])).max() > 0
    print("c_sw", hydrostatic, errs)
    assert max(errs.values()) <= TOL, errs
