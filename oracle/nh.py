"""ORACLE (test infrastructure only).  Non-hydrostatic pieces restated in torch float64:

  model/nh_utils_nlm.F90   update_dz_c :43-182, update_dz_d :183-296, Riem_Solver_c :297-401,
                           SIM1_solver :1177-1308, SIM_solver :1310-1466, edge_profile :1519-1625 (non-uniform branch)
  model/nh_core_nlm.F90    Riem_Solver3 :40-206 (a_imp > 0.999 -> SIM1_solver, use_logp = F, fp_out = F)
  model/dyn_core_nlm.F90   the non-hydrostatic branch of dyn_core :78-1040, pk3_halo :1129, pe_halo :1232
  (TL: model_tlmadm/nh_utils_tlm.F90, nh_core_tlm.F90; AD: nh_utils_adm.F90, nh_core_adm.F90)

Configuration: a_imp = 1 (fully implicit SIM1) or 0.5 < a_imp <= 0.999 (SIM_solver, Riem_Solver3 only; scale_m = 0), p_fac = 0.05, no MOIST_CAPPA / USE_COND,
beta = 0 (nh_p_grad), use_logp = F.   parity: pinned by tests/test_ref_tlm.py (SIM1_SOLVER / RIEM_SOLVER_C / RIEM_SOLVER3 /
UPDATE_DZ_C / UPDATE_DZ_D / EDGE_PROFILE _TLM) and by the reference's DYN_CORE_TLM executed on six tiles (tests/test_ref_golden.py: 5e-15);
the SIM solver (a_imp < 0.999) has no reference pin.
"""
import numpy as np
import torch
from .cubed_sphere import R, fill_4corners
from .sw_core import c_sw, S, put, Z, O, sg
from .d_sw import d_sw, del6_vt_flux
from . import tp_core as tp
from .dyn_core import halo_of, p_grad_c, grad_p, level_params, geopk, heat_update, sides, level_params_pert

DZ_MIN = 2.0
R3 = 1. / 3.


def sim1_solver(dt, dm2, pm2, pem, w1, dz2, pt2, ws, rgas, gama, kappa, p_fac):
    """SIM1_solver: lists over k of tensors.  Returns w2 (K), pe (K+1), dz2_new (K)"""
    K = len(dm2)
    t1g = gama * 2. * dt * dt
    rdt = 1. / dt
    capa1 = kappa - 1.
    pe = [torch.exp(gama * torch.log(-dm2[k] / dz2[k] * rgas * pt2[k])) - pm2[k] for k in range(K)]
    g_rat = [None] * K; bb = [None] * K; dd = [None] * K
    for k in range(K - 1):
        g_rat[k] = dm2[k] / dm2[k + 1]
        bb[k] = 2. * (1. + g_rat[k])
        dd[k] = 3. * (pe[k] + g_rat[k] * pe[k + 1])
    bb[K - 1] = 2. + 0. * dm2[0]
    dd[K - 1] = 3. * pe[K - 1]
    pp = [None] * (K + 1); gam = [None] * K
    bet = bb[0]
    pp[0] = 0. * dm2[0]
    pp[1] = dd[0] / bet
    for k in range(1, K):
        gam[k] = g_rat[k - 1] / bet
        bet = bb[k] - gam[k]
        pp[k + 1] = (dd[k] - pp[k]) / bet
    for k in range(K - 1, 0, -1):
        pp[k] = pp[k] - gam[k] * pp[k + 1]
    aa = [None] * K
    for k in range(1, K):
        aa[k] = t1g / (dz2[k - 1] + dz2[k]) * (pem[k] + pp[k])
    w2 = [None] * K
    bet = dm2[0] - aa[1]
    w2[0] = (dm2[0] * w1[0] + dt * pp[1]) / bet
    for k in range(1, K - 1):
        gam[k] = aa[k] / bet
        bet = dm2[k] - (aa[k] + aa[k + 1] + aa[k] * gam[k])
        w2[k] = (dm2[k] * w1[k] + dt * (pp[k + 1] - pp[k]) - aa[k] * w2[k - 1]) / bet
    p1 = t1g / dz2[K - 1] * (pem[K] + pp[K])
    gam[K - 1] = aa[K - 1] / bet
    bet = dm2[K - 1] - (aa[K - 1] + p1 + aa[K - 1] * gam[K - 1])
    w2[K - 1] = (dm2[K - 1] * w1[K - 1] + dt * (pp[K] - pp[K - 1]) - p1 * ws - aa[K - 1] * w2[K - 2]) / bet
    for k in range(K - 2, -1, -1):
        w2[k] = w2[k] - gam[k + 1] * w2[k + 1]
    pe2 = [None] * (K + 1)
    pe2[0] = 0. * dm2[0]
    for k in range(K):
        pe2[k + 1] = pe2[k] + dm2[k] * (w2[k] - w1[k]) * rdt
    dzn = [None] * K
    p1 = (pe2[K - 1] + 2. * pe2[K]) * R3
    dzn[K - 1] = -dm2[K - 1] * rgas * pt2[K - 1] * torch.exp(capa1 * torch.log(torch.maximum(p_fac * pm2[K - 1], p1 + pm2[K - 1])))
    for k in range(K - 2, -1, -1):
        p1 = (pe2[k] + bb[k] * pe2[k + 1] + g_rat[k] * pe2[k + 2]) * R3 - g_rat[k] * p1
        dzn[k] = -dm2[k] * rgas * pt2[k] * torch.exp(capa1 * torch.log(torch.maximum(p_fac * pm2[k], p1 + pm2[k])))
    return w2, pe2, dzn


def sim_solver(dt, dm2, pm2, pem, w1, dz2, pt2, ws, rgas, gama, kappa, p_fac, alpha, scale_m=0.):
    """SIM_solver (model/nh_utils_nlm.F90:1310-1466): off-centred (alpha = a_imp in (0.5, 0.999]) semi-implicit
    solver; lists over k of tensors.  Returns w2 (K), pe2 (K+1, blended with pp at the end), dz2_new (K)"""
    K = len(dm2)
    beta = 1. - alpha
    ra = 1. / alpha
    t2 = beta / alpha
    t1g = 2. * gama * (alpha * dt) ** 2
    rdt = 1. / dt
    capa1 = kappa - 1.
    pe = [torch.exp(gama * torch.log(-dm2[k] / dz2[k] * rgas * pt2[k])) - pm2[k] for k in range(K)]
    g_rat = [None] * K; bb = [None] * K; dd = [None] * K
    for k in range(K - 1):
        g_rat[k] = dm2[k] / dm2[k + 1]
        bb[k] = 2. * (1. + g_rat[k])
        dd[k] = 3. * (pe[k] + g_rat[k] * pe[k + 1])
    bb[K - 1] = 2. + 0. * dm2[0]
    dd[K - 1] = 3. * pe[K - 1]
    pp = [None] * (K + 1); gam = [None] * K
    bet = bb[0]
    pp[0] = 0. * dm2[0]
    pp[1] = dd[0] / bet
    for k in range(1, K):
        gam[k] = g_rat[k - 1] / bet
        bet = bb[k] - gam[k]
        pp[k + 1] = (dd[k] - pp[k]) / bet
    for k in range(K - 1, 0, -1):
        pp[k] = pp[k] - gam[k] * pp[k + 1]
    pef = [pem[k] + pp[k] for k in range(K + 1)]          # "pe2 is Full p"
    aa = [None] * K; wk = [None] * K
    for k in range(1, K):
        aa[k] = t1g / (dz2[k - 1] + dz2[k]) * pef[k]
        wk[k] = t2 * aa[k] * (w1[k - 1] - w1[k])
        aa[k] = aa[k] - scale_m * dm2[0]
    w2 = [None] * K
    bet = dm2[0] - aa[1]
    w2[0] = (dm2[0] * w1[0] + dt * pp[1] + wk[1]) / bet
    for k in range(1, K - 1):
        gam[k] = aa[k] / bet
        bet = dm2[k] - (aa[k] + aa[k + 1] + aa[k] * gam[k])
        w2[k] = (dm2[k] * w1[k] + dt * (pp[k + 1] - pp[k]) + wk[k + 1] - wk[k] - aa[k] * w2[k - 1]) / bet
    wk1 = t1g / dz2[K - 1] * pef[K]
    gam[K - 1] = aa[K - 1] / bet
    bet = dm2[K - 1] - (aa[K - 1] + wk1 + aa[K - 1] * gam[K - 1])
    w2[K - 1] = (dm2[K - 1] * w1[K - 1] + dt * (pp[K] - pp[K - 1]) - wk[K - 1] + wk1 * (t2 * w1[K - 1] - ra * ws)
                 - aa[K - 1] * w2[K - 2]) / bet
    for k in range(K - 2, -1, -1):
        w2[k] = w2[k] - gam[k + 1] * w2[k + 1]
    pe2 = [None] * (K + 1)
    pe2[0] = 0. * dm2[0]
    for k in range(K):
        pe2[k + 1] = pe2[k] + (dm2[k] * (w2[k] - w1[k]) * rdt - beta * (pp[k + 1] - pp[k])) * ra
    dzn = [None] * K
    p1 = (pe2[K - 1] + 2. * pe2[K]) * R3
    dzn[K - 1] = -dm2[K - 1] * rgas * pt2[K - 1] * torch.exp(capa1 * torch.log(torch.maximum(p_fac * pm2[K - 1], p1 + pm2[K - 1])))
    for k in range(K - 2, -1, -1):
        p1 = (pe2[k] + bb[k] * pe2[k + 1] + g_rat[k] * pe2[k + 2]) * R3 - g_rat[k] * p1
        dzn[k] = -dm2[k] * rgas * pt2[k] * torch.exp(capa1 * torch.log(torch.maximum(p_fac * pm2[k], p1 + pm2[k])))
    pe2 = [pe2[k] + beta * (pp[k] - pe2[k]) for k in range(K + 1)]
    return w2, pe2, dzn


def riem_solver_c(dt, delp, pt, gz, w3, ws, hs, cfg):
    """Riem_Solver_c: returns pef (full pressure) and the new gz; [6,K(+1),..] arrays (all points)"""
    K = delp.shape[1]
    akap, ptop, rdgas, grav, p_fac = cfg["akap"], cfg["ptop"], cfg["rdgas"], cfg["grav"], cfg["p_fac"]
    gama = 1. / (1. - akap); rgrav = 1. / grav
    z1 = torch.zeros_like(delp[:, :1])
    pem = torch.cumsum(torch.cat([z1 + ptop, delp], dim=1), dim=1)
    dz2 = gz[:, 1:] - gz[:, :-1]
    pm2 = delp / torch.log(pem[:, 1:] / pem[:, :-1])
    dm = delp * rgrav
    L = lambda a: list(torch.unbind(a, dim=1))
    w2, pe2, dzn = sim1_solver(dt, L(dm), L(pm2), L(pem), L(w3), L(dz2), L(pt), ws[:, 0], rdgas, gama, akap, p_fac)
    pef = torch.stack([pem[:, 0]] + [pe2[k] + pem[:, k] for k in range(1, K + 1)], dim=1)
    gzl = [None] * (K + 1)
    gzl[K] = hs[:, 0] + 0. * dzn[0]
    for k in range(K - 1, -1, -1):
        gzl[k] = gzl[k + 1] - dzn[k] * grav
    return pef, torch.stack(gzl, dim=1)


def riem_solver3(dt, delp, pt, zh, w, ws, zs, cfg):
    """Riem_Solver3: returns w, delz, zh, ppe (perturbation pressure)"""
    K = delp.shape[1]
    akap, ptop, rdgas, grav, p_fac = cfg["akap"], cfg["ptop"], cfg["rdgas"], cfg["grav"], cfg["p_fac"]
    gama = 1. / (1. - akap); rgrav = 1. / grav
    z1 = torch.zeros_like(delp[:, :1])
    pem = torch.cumsum(torch.cat([z1 + ptop, delp], dim=1), dim=1)
    peln2 = torch.log(pem)
    pm2 = delp / (peln2[:, 1:] - peln2[:, :-1])
    dm = delp * rgrav
    dz2 = zh[:, 1:] - zh[:, :-1]
    L = lambda a: list(torch.unbind(a, dim=1))
    a_imp = cfg.get("a_imp", 1.0)
    if a_imp > 0.999:
        w2, pe2, dzn = sim1_solver(dt, L(dm), L(pm2), L(pem), L(w), L(dz2), L(pt), ws[:, 0], rdgas, gama, akap, p_fac)
    elif a_imp > 0.5:       # model/nh_core_nlm.F90:136-152
        w2, pe2, dzn = sim_solver(dt, L(dm), L(pm2), L(pem), L(w), L(dz2), L(pt), ws[:, 0], rdgas, gama, akap, p_fac, a_imp)
    else:
        raise NotImplementedError("a_imp <= 0.5 (RIM_2D / SIM3 solvers)")
    zl = [None] * (K + 1)
    zl[K] = zs[:, 0] + 0. * dzn[0]
    for k in range(K - 1, -1, -1):
        zl[k] = zl[k + 1] - dzn[k]
    return torch.stack(w2, dim=1), torch.stack(dzn, dim=1), torch.stack(zl, dim=1), torch.stack(pe2, dim=1)


def edge_profile(q, dp0):
    """edge_profile (non-uniform grid, limiter = 0): layer means [6,K,..] -> interface values [6,K+1,..]"""
    K = q.shape[1]
    Q = lambda k: q[:, k]
    g0 = dp0[1] / dp0[0]
    xt1 = 2. * g0 * (g0 + 1.)
    bet = g0 * (g0 + 0.5)
    qe = [None] * (K + 1); gam = [None] * (K + 1)
    qe[0] = (xt1 * Q(0) + Q(1)) / bet
    gam[0] = (1. + g0 * (g0 + 1.5)) / bet
    gk = None
    for k in range(1, K):
        gk = dp0[k - 1] / dp0[k]
        bet = 2. + 2. * gk - gam[k - 1]
        qe[k] = (3. * (Q(k - 1) + gk * Q(k)) - qe[k - 1]) / bet
        gam[k] = gk / bet
    a_bot = 1. + gk * (gk + 1.5)
    xt1 = 2. * gk * (gk + 1.)
    xt2 = gk * (gk + 0.5) - a_bot * gam[K - 1]
    qe[K] = (xt1 * Q(K - 1) + Q(K - 2) - a_bot * qe[K - 1]) / xt2
    for k in range(K - 1, -1, -1):
        qe[k] = qe[k] - gam[k] * qe[k + 1]
    return torch.stack(qe, dim=1)


def update_dz_c(dt, dp0, zs, ut, vt, gz, g):
    """update_dz_c: returns gz (valid is-1..ie+1) and ws [6,1,..]"""
    N, npx, npy = g.N, g.npx, g.npy
    is_, ie, js, je = 1, N, 1, N
    K = ut.shape[1]
    rdt = 1. / dt
    top_ratio = dp0[0] / (dp0[0] + dp0[1])
    bot_ratio = dp0[K - 1] / (dp0[K - 2] + dp0[K - 1])
    def interface(a):
        lev = [a[:, 0] + (a[:, 0] - a[:, 1]) * top_ratio]
        for k in range(1, K):
            ir = 1. / (dp0[k - 1] + dp0[k])
            lev.append((dp0[k] * a[:, k - 1] + dp0[k - 1] * a[:, k]) * ir)
        lev.append(a[:, K - 1] + (a[:, K - 1] - a[:, K - 2]) * bot_ratio)
        return torch.stack(lev, dim=1)
    xfx = interface(ut); yfx = interface(vt)
    gz1 = fill_4corners(gz, npx, npy, 1)
    i0, i1, j0, j1 = is_ - 1, ie + 2, js - 1, je + 1
    xf = S(xfx, i0, i1, j0, j1)
    fx = xf * torch.where(xf > 0., S(gz1, i0 - 1, i1 - 1, j0, j1), S(gz1, i0, i1, j0, j1))
    gz2 = fill_4corners(gz1, npx, npy, 2)
    i0, i1, j0, j1 = is_ - 1, ie + 1, js - 1, je + 2
    yf = S(yfx, i0, i1, j0, j1)
    fy = yf * torch.where(yf > 0., S(gz2, i0, i1, j0 - 1, j1 - 1), S(gz2, i0, i1, j0, j1))
    i0, i1, j0, j1 = is_ - 1, ie + 1, js - 1, je + 1
    ar = S(g.area, i0, i1, j0, j1)
    gzn = (S(gz2, i0, i1, j0, j1) * ar + (fx[..., :, :-1] - fx[..., :, 1:]) + (fy[..., :-1, :] - fy[..., 1:, :])) / \
        (ar + (xf[..., :, :-1] - xf[..., :, 1:]) + (yf[..., :-1, :] - yf[..., 1:, :]))
    ws = (S(zs, i0, i1, j0, j1)[:, 0] - gzn[:, K]) * rdt
    lev = [None] * (K + 1)
    lev[K] = gzn[:, K]
    for k in range(K - 1, -1, -1):
        lev[k] = torch.maximum(gzn[:, k], lev[k + 1] + DZ_MIN)
    gz_out = put(gz2, i0, i1, j0, j1, torch.stack(lev, dim=1))
    ws_out = put(torch.zeros_like(zs), i0, i1, j0, j1, ws[:, None])
    return gz_out, ws_out


def update_dz_d(ndif, damp, hord, dp0, zs, zh, crx, cry, xfx, yfx, g, rdt, hord_pert=None):
    """update_dz_d: ndif, damp per-level lists (length K; K+1-th = K-th).  Returns zh, ws"""
    N, npx, npy = g.N, g.npx, g.npy
    is_, ie, js, je = 1, N, 1, N
    isd, ied, jsd, jed = g.isd, g.ied, g.jsd, g.jed
    K = crx.shape[1]
    ndif = list(ndif) + [ndif[-1]]; damp = list(damp) + [damp[-1]]
    crx_a = edge_profile(crx, dp0); xfx_a = edge_profile(xfx, dp0)
    cry_a = edge_profile(cry, dp0); yfx_a = edge_profile(yfx, dp0)
    ra_x = put(Z(zh), is_, ie, jsd, jed, S(g.area, is_, ie, jsd, jed) + (S(xfx_a, is_, ie, jsd, jed) - S(xfx_a, is_ + 1, ie + 1, jsd, jed)))
    ra_y = put(Z(zh), isd, ied, js, je, S(g.area, isd, ied, js, je) + (S(yfx_a, isd, ied, js, je) - S(yfx_a, isd, ied, js + 1, je + 1)))
    hl = [hord] * (K + 1) if isinstance(hord, int) else list(hord) + [hord[-1]]
    fx, fy, z2 = tp.fv_tp_2d(zh, crx_a, cry_a, hl if len(set(hl)) > 1 else hl[0], xfx_a, yfx_a, g, ra_x, ra_y)
    if hord_pert is not None and hord_pert != hord:      # model_tlmadm/nh_utils_tlm.F90:496-560
        from .d_sw import splice
        fxp, fyp, _ = tp.fv_tp_2d(zh, crx_a, cry_a, hord_pert, xfx_a, yfx_a, g, ra_x, ra_y)
        fx, fy = splice(fxp, fx), splice(fyp, fy)
    C = (is_, ie, js, je)
    base = (S(zh, *C) * S(g.area, *C) + (S(fx, *C) - S(fx, is_ + 1, ie + 1, js, je)) + (S(fy, *C) - S(fy, is_, ie, js + 1, je + 1))) / \
        ((S(ra_x, *C) + S(ra_y, *C)) - S(g.area, *C))
    if any(d > 1.e-5 for d in damp):
        keys = [(n, d) if d > 1.e-5 else None for n, d in zip(ndif, damp)]
        var = {}
        for key in set(keys):
            if key is None:
                var["None"] = torch.zeros_like(base)
            else:
                fx2, fy2 = del6_vt_flux(key[0], key[1], z2, g)
                var[str(key)] = ((S(fx2, *C) - S(fx2, is_ + 1, ie + 1, js, je)) + (S(fy2, *C) - S(fy2, is_, ie, js + 1, je + 1))) * S(g.rarea, *C)
        base = base + tp.lev_select([str(k) for k in keys], var)
    ws = (S(zs, *C)[:, 0] - base[:, K]) * rdt
    lev = [None] * (K + 1)
    lev[K] = base[:, K]
    for k in range(K - 1, -1, -1):
        lev[k] = torch.maximum(base[:, k], lev[k + 1] + DZ_MIN)
    return put(zh, *C, torch.stack(lev, dim=1)), put(torch.zeros_like(zs), *C, ws[:, None])


def dyn_core_nh(st, g, cfg, ak, bk, first_call=True):
    """non-hydrostatic acoustic loop (dyn_core_nlm.F90, .not. hydrostatic branch).
    st: u v w delz pt delp phis (halos of u, v, pt, delp valid).  Returns u v w delz pt delp
    mfx mfy cx cy pe pk peln pkz ws."""
    N = g.N
    is_, ie, js, je = 1, N, 1, N
    isd, ied, jsd, jed = g.isd, g.ied, g.jsd, g.jed
    halo, getb = halo_of(N)
    K = st["delp"].shape[1]
    n_split = cfg["n_split"]
    dt = cfg["bdt"] / n_split
    dt2 = 0.5 * dt
    rdt = 1. / dt
    grav = cfg["grav"]; ptop, akap = cfg["ptop"], cfg["akap"]
    cfg, cfgp = sides(cfg)
    prm = level_params(cfg, K)
    prm["hydrostatic"] = False
    pp = level_params_pert(cfgp, K, prm) if cfgp else None
    dp_ref = [(ak[k + 1] - ak[k]) + (bk[k + 1] - bk[k]) * 1.e5 for k in range(K)]
    u, v, w, delz, pt, delp = st["u"], st["v"], st["w"], st["delz"], st["pt"], st["delp"]
    hs = st["phis"]
    zs = hs / grav
    mfx = Z(u); mfy = Z(u); cx = Z(u); cy = Z(u)
    zh = None
    heat = None
    C = (slice(None), slice(None), R(js, je), R(is_, ie))
    for it in range(1, n_split + 1):
        w = halo.scalar(w)
        if it == 1:
            lev = [None] * (K + 1)
            lev[K] = zs[:, 0] + 0. * delz[:, 0]
            for k in range(K - 1, -1, -1):
                lev[k] = lev[k + 1] - delz[:, k]
            gzf = torch.stack(lev, dim=1)
            gz = torch.zeros_like(gzf); gz[C] = gzf[C]
            gz = halo.scalar(gz)
            zh = gz
        else:
            gz = zh
        c = c_sw(delp, pt, u, v, w, g, dt2, False, cfg["nord"])
        divgd = halo.corner(c["divg_d"]) if cfg["nord"] > 0 else c["divg_d"]
        gz, ws3 = update_dz_c(dt2, dp_ref, zs, c["ut"], c["vt"], gz, g)
        # solve only where the reference does (is-1..ie+1): elsewhere the inputs are undefined
        C1 = (slice(None), slice(None), R(js - 1, je + 1), R(is_ - 1, ie + 1))
        pef1, gz1 = riem_solver_c(dt2, c["delpc"][C1], c["ptc"][C1], gz[C1], c["wc"][C1], ws3[C1], hs[C1], cfg)
        pkc = torch.zeros_like(gz); pkc[C1] = pef1
        gzn = torch.zeros_like(gz); gzn[C1] = gz1
        gz = gzn
        uc, vc = p_grad_c(dt2, c["delpc"], pkc, gz, c["uc"], c["vc"], g, False)
        uc, vc = halo.cgrid(uc, vc)
        d = d_sw(delp, pt, u, v, w, uc, vc, c["ua"], c["va"], divgd, g, dt, prm, pp)
        mfx = put(mfx, is_, ie + 1, js, je, S(mfx, is_, ie + 1, js, je) + S(d["fx"], is_, ie + 1, js, je))
        mfy = put(mfy, is_, ie, js, je + 1, S(mfy, is_, ie, js, je + 1) + S(d["fy"], is_, ie, js, je + 1))
        cx = put(cx, is_, ie + 1, jsd, jed, S(cx, is_, ie + 1, jsd, jed) + S(d["crx"], is_, ie + 1, jsd, jed))
        cy = put(cy, isd, ied, js, je + 1, S(cy, isd, ied, js, je + 1) + S(d["cry"], isd, ied, js, je + 1))
        if "heat" in d:
            heat = d["heat"] if heat is None else heat + d["heat"]
        delp = halo.scalar(d["delp"]); pt = halo.scalar(d["pt"]); w = d["w"]
        zh, ws = update_dz_d(prm["nord_v"], prm["damp_v"], cfg["hord_tm"], dp_ref, zs, zh, d["crx"], d["cry"], d["xfx"], d["yfx"], g, rdt,
                             hord_pert=cfgp["hord_tm"] if cfgp else None)
        wn, dzn, zhn, ppe = riem_solver3(dt, delp[C], pt[C], zh[C], w[C], ws[C], zs[C], cfg)
        w = torch.zeros_like(w); w[C] = wn
        delz = torch.zeros_like(delz); delz[C] = dzn
        zhc = torch.zeros_like(zh); zhc[C] = zhn
        pc = torch.zeros_like(zh); pc[C] = ppe
        zh = halo.scalar(zhc); pkc = halo.scalar(pc)
        # pk3 (interior from the solver, halo ring from pk3_halo) = exp(akap log(ptop + cumsum(delp))) everywhere
        pk3, _, pe, peln, _ = geopk(delp, pt, hs, g, ptop, akap, cfg["cp_air"], 2, True)
        gz = zh * grav
        if cfg.get("beta", 0.0) > 0.0:      # split_p_grad (dyn_core_nlm.F90:874-875), beta_d = 0 on the first sub-step (:373-375)
            u, v, du_dv = grad_p(d["u"], d["v"], pk3, gz, g, dt, ptop ** akap, pp=pkc, delp=delp, beta=0.0 if it == 1 else cfg["beta"],
                                 du_dv=None if it == 1 else du_dv)
        else:
            u, v = grad_p(d["u"], d["v"], pk3, gz, g, dt, ptop ** akap, pp=pkc, delp=delp)
        if it == n_split:
            u, v = getb(u, v)
        else:
            u, v = halo.dgrid(u, v)
    if heat is not None:
        pt = heat_update(heat, pt, delp, delz, g, cfg, halo, False)
    return dict(u=u, v=v, w=w, delz=delz, pt=pt, delp=delp, mfx=mfx, mfy=mfy, cx=cx, cy=cy, pe=pe, pk=pk3, peln=peln,
                pkz=torch.zeros_like(delp), ws=ws)
