"""ORACLE (test infrastructure only).  dyn_core: the acoustic sub-cycle, restated in torch
float64 from model/dyn_core_nlm.F90  dyn_core :78-1040, p_grad_c :1369, nh_p_grad :1431,
one_grad_p :1645, geopk :1954 (TL model_tlmadm/dyn_core_tlm.F90:93, AD dyn_core_adm.F90:115/1686).

Configuration restated: grid_type 0, non-nested, beta = 0, d_ext = 0 (divg2 = 0), d_con = 0,
inline_q = F, no USE_COND.  Halo exchanges (mpp_update_domains, un-vendored FMS) are the
index maps of oracle/cubed_sphere.py.

parity: P_GRAD_C_TLM / NH_P_GRAD_TLM pinned by tests/test_ref_tlm.py, the non-hydrostatic loop by the reference's DYN_CORE_TLM executed
(tests/test_ref_golden.py); the hydrostatic loop, del2_cubed and the heating update have no reference pin.
"""
import numpy as np
import torch
from .cubed_sphere import R, NG, Halo, neighbour, to_neighbour, STAG, copy_corners
from .sw_core import c_sw, S, put, Z, O
from .sw_core import P as Pt
from .d_sw import d_sw
from .a2b_edge import a2b_ord4, a2b_ord2


class GetBoundary:
    """mpp_get_boundary(u, v, gridtype=DGRID_NE) for whole tiles (dyn_core_nlm.F90:943-955):
    north row of u and east column of v come from the tile owning them as south/west edge."""

    def __init__(self, N):
        rec = []
        o = NG - 1
        st = ("ystag", "xstag")
        for t in range(6):
            for e in (0, 1):
                for n in range(1, N + 1):
                    i, j = (n, N + 1) if e == 0 else (N + 1, n)
                    ox, oy = STAG[st[e]]
                    x, y = i - 1 + ox, j - 1 + oy
                    side = "N" if e == 0 else "E"
                    tb, kind = neighbour(t, side)
                    xb, yb, rot = to_neighbour(kind, x, y, N)
                    sc, sg = e, 1.0
                    if rot != 0:
                        sc = 1 - e
                        sg = (-1.0 if e == 0 else 1.0) if rot == 1 else (1.0 if e == 0 else -1.0)
                    sox, soy = STAG[st[sc]]
                    si, sj = int(round(xb - sox + 1)), int(round(yb - soy + 1))
                    rec.append((t, j + o, i + o, e, tb, sj + o, si + o, sc, sg))
        self.rec = rec

    def __call__(self, u, v):
        un, vn = u.clone(), v.clone()
        src = (u, v)
        for e, dst in ((0, un), (1, vn)):
            for sc in (0, 1):
                r = [x for x in self.rec if x[3] == e and x[7] == sc]
                if not r:
                    continue
                dt = torch.tensor([x[0] for x in r]); dj = torch.tensor([x[1] for x in r]); di = torch.tensor([x[2] for x in r])
                st = torch.tensor([x[4] for x in r]); sj = torch.tensor([x[5] for x in r]); si = torch.tensor([x[6] for x in r])
                sg = torch.tensor([x[8] for x in r], dtype=u.dtype)
                dst[dt, :, dj, di] = sg[:, None] * src[sc][st, :, sj, si]
        return un, vn


_cache = {}


def halo_of(N):
    if N not in _cache:
        _cache[N] = (Halo(N), GetBoundary(N))
    return _cache[N]


def geopk(delp, pt, hs, g, ptop, akap, cp_air, halo, cg):
    """dyn_core_nlm.F90:1954-2087.  returns pk, gz, pe, peln [6,K+1,..] and pkz [6,K,..]"""
    K = delp.shape[1]
    z1 = torch.zeros_like(delp[:, :1])
    pe = torch.cumsum(torch.cat([z1 + ptop, delp], dim=1), dim=1)      # same summation order as the column loop
    peln = torch.log(pe)
    pk = torch.exp(akap * peln)
    pk = torch.cat([z1 + ptop ** akap, pk[:, 1:]], dim=1)
    peln = torch.cat([z1 + np.log(ptop), peln[:, 1:]], dim=1)
    dpk = pk[:, 1:] - pk[:, :-1]
    incr = cp_air * pt * dpk
    seq = torch.cat([hs.expand_as(z1) + 0.0 * z1, torch.flip(incr, dims=[1])], dim=1)   # hs, incr(K), incr(K-1), ...
    gz = torch.flip(torch.cumsum(seq, dim=1), dims=[1])
    pkz = dpk / (akap * (peln[:, 1:] - peln[:, :-1]))
    return pk, gz, pe, peln, pkz


def p_grad_c(dt2, delpc, pkc, gz, uc, vc, g, hydrostatic):
    """dyn_core_nlm.F90:1369-1429"""
    N = g.N
    is_, ie, js, je = 1, N, 1, N
    lo = lambda a: a[:, :-1]
    hi = lambda a: a[:, 1:]
    wk = (hi(pkc) - lo(pkc)) if hydrostatic else delpc
    i0, i1, j0, j1 = is_, ie + 1, js, je
    ucn = S(uc, i0, i1, j0, j1) + dt2 * S(g.rdxc, i0, i1, j0, j1) / (S(wk, i0 - 1, i1 - 1, j0, j1) + S(wk, i0, i1, j0, j1)) * (
        (S(hi(gz), i0 - 1, i1 - 1, j0, j1) - S(lo(gz), i0, i1, j0, j1)) * (S(hi(pkc), i0, i1, j0, j1) - S(lo(pkc), i0 - 1, i1 - 1, j0, j1)) +
        (S(lo(gz), i0 - 1, i1 - 1, j0, j1) - S(hi(gz), i0, i1, j0, j1)) * (S(hi(pkc), i0 - 1, i1 - 1, j0, j1) - S(lo(pkc), i0, i1, j0, j1)))
    i0, i1, j0, j1 = is_, ie, js, je + 1
    vcn = S(vc, i0, i1, j0, j1) + dt2 * S(g.rdyc, i0, i1, j0, j1) / (S(wk, i0, i1, j0 - 1, j1 - 1) + S(wk, i0, i1, j0, j1)) * (
        (S(hi(gz), i0, i1, j0 - 1, j1 - 1) - S(lo(gz), i0, i1, j0, j1)) * (S(hi(pkc), i0, i1, j0, j1) - S(lo(pkc), i0, i1, j0 - 1, j1 - 1)) +
        (S(lo(gz), i0, i1, j0 - 1, j1 - 1) - S(hi(gz), i0, i1, j0, j1)) * (S(hi(pkc), i0, i1, j0 - 1, j1 - 1) - S(lo(pkc), i0, i1, j0, j1)))
    return put(uc, is_, ie + 1, js, je, ucn), put(vc, is_, ie, js, je + 1, vcn)


def grad_p(u, v, pk, gz, g, dt, top_value, pp=None, delp=None, beta=None, du_dv=None, divg2=None):
    """one_grad_p (:1645, hydrostatic, d_ext = 0) when pp is None, nh_p_grad (:1431) otherwise.
    beta is not None (the caller's beta_d, 0 on the first acoustic sub-step, dyn_core :373-375): grad1_p_update (:1781-1872, d_ext = 0 so
    divg2 = 0 :726) / split_p_grad (:1531-1643): u += beta * du_prev, then the hydrostatic part enters with alpha = 1 - beta and is
    handed to the next sub-step; du_dv = (du_prev, dv_prev) on the output rectangles or None; returns (u, v, (du, dv)).
    divg2 (hydrostatic, d_ext > 0; [6, 1, NY, NX] at the corners): the external-mode damping term of one_grad_p (:1713-1727, :1758-1771)
    and grad1_p_update (:1858, :1867)."""
    N = g.N
    is_, ie, js, je = 1, N, 1, N
    pkb = a2b_ord4(pk, g)
    pkb = torch.cat([torch.zeros_like(pkb[:, :1]) + top_value, pkb[:, 1:]], dim=1)
    gzb = a2b_ord4(gz, g)
    lo = lambda a: a[:, :-1]
    hi = lambda a: a[:, 1:]
    wk = hi(pkb) - lo(pkb)

    def lin(q, i0, i1, j0, j1, di, dj):
        # ((gz(k+1) - gz'(k)) * (q'(k+1) - q(k)) + (gz(k) - gz'(k+1)) * (q(k+1) - q'(k)))  with ' = neighbour (di,dj)
        return ((S(hi(gzb), i0, i1, j0, j1) - S(lo(gzb), i0 + di, i1 + di, j0 + dj, j1 + dj)) * (S(hi(q), i0 + di, i1 + di, j0 + dj, j1 + dj) - S(lo(q), i0, i1, j0, j1)) +
                (S(lo(gzb), i0, i1, j0, j1) - S(hi(gzb), i0 + di, i1 + di, j0 + dj, j1 + dj)) * (S(hi(q), i0, i1, j0, j1) - S(lo(q), i0 + di, i1 + di, j0 + dj, j1 + dj)))
    i0, i1, j0, j1 = is_, ie, js, je + 1
    du = dt / (S(wk, i0, i1, j0, j1) + S(wk, i0 + 1, i1 + 1, j0, j1)) * lin(pkb, i0, i1, j0, j1, 1, 0)
    if beta is not None:
        alpha = 1.0 - beta
        ub = S(u, i0, i1, j0, j1) + (beta * du_dv[0] if du_dv is not None else 0.0)
        if divg2 is not None:
            ub = (ub + S(divg2, i0, i1, j0, j1)) - S(divg2, i0 + 1, i1 + 1, j0, j1)
    if beta is not None and pp is None:
        un = (ub + alpha * du) * S(g.rdx, i0, i1, j0, j1)
    elif pp is None:
        wk2 = 0.0 if divg2 is None else S(divg2, i0, i1, j0, j1) - S(divg2, i0 + 1, i1 + 1, j0, j1)
        un = S(g.rdx, i0, i1, j0, j1) * (wk2 + S(u, i0, i1, j0, j1) + du)
    else:
        ppb = a2b_ord4(pp, g)
        ppb = torch.cat([torch.zeros_like(ppb[:, :1]), ppb[:, 1:]], dim=1)
        wk1 = a2b_ord4(delp, g)
        dn = dt / (S(wk1, i0, i1, j0, j1) + S(wk1, i0 + 1, i1 + 1, j0, j1)) * lin(ppb, i0, i1, j0, j1, 1, 0)
        un = ((S(u, i0, i1, j0, j1) + du + dn) if beta is None else (ub + alpha * du + dn)) * S(g.rdx, i0, i1, j0, j1)
    i0, i1, j0, j1 = is_, ie + 1, js, je
    dv = dt / (S(wk, i0, i1, j0, j1) + S(wk, i0, i1, j0 + 1, j1 + 1)) * lin(pkb, i0, i1, j0, j1, 0, 1)
    if beta is not None:
        vb = S(v, i0, i1, j0, j1) + (beta * du_dv[1] if du_dv is not None else 0.0)
        if divg2 is not None:
            vb = (vb + S(divg2, i0, i1, j0, j1)) - S(divg2, i0, i1, j0 + 1, j1 + 1)
    if beta is not None and pp is None:
        vn = (vb + alpha * dv) * S(g.rdy, i0, i1, j0, j1)
    elif pp is None:
        wk1 = 0.0 if divg2 is None else S(divg2, i0, i1, j0, j1) - S(divg2, i0, i1, j0 + 1, j1 + 1)
        vn = S(g.rdy, i0, i1, j0, j1) * (wk1 + S(v, i0, i1, j0, j1) + dv)
    else:
        dn = dt / (S(wk1, i0, i1, j0, j1) + S(wk1, i0, i1, j0 + 1, j1 + 1)) * lin(ppb, i0, i1, j0, j1, 0, 1)
        vn = ((S(v, i0, i1, j0, j1) + dv + dn) if beta is None else (vb + alpha * dv + dn)) * S(g.rdy, i0, i1, j0, j1)
    if beta is not None:
        return put(u, is_, ie, js, je + 1, un), put(v, is_, ie + 1, js, je, vn), (du, dv)
    return put(u, is_, ie, js, je + 1, un), put(v, is_, ie + 1, js, je, vn)


def level_params(cfg, K):
    """per-level switches of the k-loop before d_sw (dyn_core_nlm.F90:579-625).  cfg: dict with
    hord_mt hord_vt hord_tm hord_dp nord d2_bg d2_bg_k1 d2_bg_k2 n_sponge vtdm4 do_vort_damp dddmp d4_bg
    (the TL module additionally uses first-order transport in the sponge, hord_*_ks = 1,
    model_tlmadm/dyn_core_tlm.F90:740-926: n_sponge_ord layers)"""
    p = {k: [] for k in ("hord_mt", "hord_vt", "hord_tm", "hord_dp", "nord", "nord_v", "nord_w", "nord_t", "d2_bg", "damp_v", "damp_w", "damp_t", "d_con")}
    ns = cfg.get("n_sponge", 0)
    for k in range(1, K + 1):
        nord_k = cfg["nord"]
        nord_v = min(2, cfg["nord"])
        d2 = min(0.20, cfg["d2_bg"])
        damp_vt = cfg["vtdm4"] if cfg["do_vort_damp"] else 0.0
        nord_w = nord_v; nord_t = nord_v; damp_w = damp_vt; damp_t = damp_vt
        d_con_k = cfg.get("d_con", 0.0)
        if K == 1 or ns < 0:
            d2 = cfg["d2_bg"]
        else:
            if k == 1:
                nord_k = 0; d2 = max(0.01, cfg["d2_bg"], cfg["d2_bg_k1"]); nord_w = 0; damp_w = d2
                if cfg["do_vort_damp"]:
                    nord_v = 0; damp_vt = 0.5 * d2
                d_con_k = 0.0
            elif k == max(2, ns - 1) and cfg["d2_bg_k2"] > 0.01:
                nord_k = 0; d2 = max(cfg["d2_bg"], cfg["d2_bg_k2"]); nord_w = 0; damp_w = d2
                if cfg["do_vort_damp"]:
                    nord_v = 0; damp_vt = 0.5 * d2
                d_con_k = 0.0
            elif k == max(3, ns) and cfg["d2_bg_k2"] > 0.05:
                nord_k = 0; d2 = max(cfg["d2_bg"], 0.2 * cfg["d2_bg_k2"]); nord_w = 0; damp_w = d2
                d_con_k = 0.0
        ho = 1 if (k <= cfg.get("n_sponge_ord", 0) or k <= cfg.get("n_sponge_ks", 0)) else None
        for n in ("hord_mt", "hord_vt", "hord_tm", "hord_dp"):
            p[n].append(ho if ho else cfg[n])
        p["nord"].append(nord_k); p["nord_v"].append(nord_v); p["nord_w"].append(nord_w); p["nord_t"].append(nord_t)
        p["d2_bg"].append(d2); p["damp_v"].append(damp_vt); p["damp_w"].append(damp_w); p["damp_t"].append(damp_t)
        p["d_con"].append(d_con_k)
    p["dddmp"] = cfg["dddmp"]; p["d4_bg"] = cfg["d4_bg"]
    if not cfg.get("d_con", 0.0) > 1.e-5:
        del p["d_con"]
    return p


def sides(cfg):
    """two-sided mode (cfg["traj"] = dict of the nonlinear model's switches; the top-level ones are then the perturbation
    model's, model_tlmadm/fv_arrays_tlmadm.F90:37-92).  Returns (trajectory cfg, perturbation cfg or None)."""
    if not cfg.get("traj"):
        return cfg, None
    ct = dict(cfg); ct.update(cfg["traj"]); ct["traj"] = None
    if cfg.get("hord_ks_traj", True):
        ct["n_sponge_ks"] = cfg.get("n_sponge", 0) - 1       # hord_*_ks_traj = 1 in the top n_sponge_pert - 1 layers
    ct["n_sponge_ord"] = 0
    return ct, cfg


def level_params_pert(cfg, K, prm_traj):
    """perturbation-side per-level switches (model_tlmadm/dyn_core_tlm.F90:835-921); cfg: the perturbation model's flags"""
    p = {k: [] for k in ("hord_mt", "hord_vt", "hord_tm", "hord_dp", "nord", "nord_v", "nord_w", "nord_t", "d2_bg", "damp_v", "damp_w", "damp_t")}
    ns = cfg.get("n_sponge", 0)
    for k in range(1, K + 1):
        ho = {n: cfg[n] for n in ("hord_mt", "hord_vt", "hord_tm", "hord_dp")}
        nord_k = cfg["nord"]; nord_v = min(2, cfg["nord"])
        d2 = min(0.20, cfg["d2_bg"])
        damp_vt = cfg["vtdm4"] if cfg["do_vort_damp"] else 0.0
        nord_w = nord_v; nord_t = nord_v; damp_w = damp_vt; damp_t = damp_vt
        if k <= ns:
            if k <= ns - 1 and cfg.get("hord_ks_pert", True):
                ho = {n: 1 for n in ho}
            nord_k = 0
            dk = cfg["d2_bg_k1"] if k == 1 else cfg["d2_bg_k2"] if k == 2 else cfg.get("d2_bg_ks", 2.0)
            d2 = max(0.01, cfg["d2_bg"], dk)
            nord_w = 0; damp_w = d2
            if cfg["do_vort_damp"]:
                nord_v = 0; damp_vt = 0.5 * d2
        for n in ho:
            p[n].append(ho[n])
        p["nord"].append(nord_k); p["nord_v"].append(nord_v); p["nord_w"].append(nord_w); p["nord_t"].append(nord_t)
        p["d2_bg"].append(d2); p["damp_v"].append(damp_vt); p["damp_w"].append(damp_w); p["damp_t"].append(damp_t)
    p["dddmp"] = cfg["dddmp"]; p["d4_bg"] = cfg["d4_bg"]; p["split_damp"] = bool(cfg.get("split_damp", False))
    for n in ("hydrostatic", "d_con"):
        if n in prm_traj:
            p[n] = prm_traj[n]
    return p


def del2_cubed(q, cd, g, nmax, halo):
    """model/dyn_core_nlm.F90:2090-2199.  q: [6,K,..] full array (halo exchanged here); returns the filtered array
    (valid on the compute domain)."""
    N, npx, npy = g.N, g.npx, g.npy
    is_, ie, js, je = 1, N, 1, N
    ntimes = min(3, nmax)
    q = halo.scalar(q)
    for n in range(1, ntimes + 1):
        nt = ntimes - n
        q = q.clone()
        for (a, b, c) in (((1, 1), (0, 1), (1, 0)), ((ie, 1), (npx, 1), (ie, 0)), ((ie, je), (npx, je), (ie, npy)), ((1, je), (0, je), (1, npy))):
            m = (Pt(q, *a) + Pt(q, *b) + Pt(q, *c)) * (1. / 3.)
            for (i, j) in (a, b, c):
                q[..., j + O, i + O] = m
        if nt > 0:
            q = copy_corners(q, npx, npy, 1)
        i0, i1, j0, j1 = is_ - nt, ie + 1 + nt, js - nt, je + nt
        fx = put(Z(q), i0, i1, j0, j1, S(g.del6_v, i0, i1, j0, j1) * (S(q, i0 - 1, i1 - 1, j0, j1) - S(q, i0, i1, j0, j1)))
        if nt > 0:
            q = copy_corners(q, npx, npy, 2)
        i0, i1, j0, j1 = is_ - nt, ie + nt, js - nt, je + 1 + nt
        fy = put(Z(q), i0, i1, j0, j1, S(g.del6_u, i0, i1, j0, j1) * (S(q, i0, i1, j0 - 1, j1 - 1) - S(q, i0, i1, j0, j1)))
        i0, i1, j0, j1 = is_ - nt, ie + nt, js - nt, je + nt
        q = put(q, i0, i1, j0, j1, S(q, i0, i1, j0, j1) + cd * S(g.rarea, i0, i1, j0, j1) *
                (S(fx, i0, i1, j0, j1) - S(fx, i0 + 1, i1 + 1, j0, j1) + S(fy, i0, i1, j0, j1) - S(fy, i0, i1, j0 + 1, j1 + 1)))
    return q


def heat_update(heat, pt, delp, aux, g, cfg, halo, hydrostatic):
    """end of dyn_core (model/dyn_core_nlm.F90:1052-1099): filter the accumulated heat source and add it to pt on the top
    n_con layers.  aux = pkz (hydrostatic) or delz.  convert_ke = F, delt_max = 1."""
    N = g.N
    K = pt.shape[1]
    if cfg["vtdm4"] > 1.e-4:
        n_con = K
    elif cfg["d2_bg_k1"] < 1.e-3:
        n_con = 0
    else:
        n_con = 1 if cfg["d2_bg_k2"] < 1.e-3 else 2
    if n_con == 0:
        return pt
    n_con = min(n_con, K)
    hs = del2_cubed(heat, 0.20 * g.da_min, g, min(3, cfg["nord"] + 1), halo)
    C = (1, N, 1, N)
    delt = abs(cfg["bdt"] * 1.0)
    ptc, hsc, dpc, ax = S(pt, *C), S(hs, *C), S(delp, *C), S(aux, *C)
    if hydrostatic:
        dtmp = hsc / (cfg["cp_air"] * dpc)
        lim = torch.where(dtmp >= 0., torch.minimum(torch.full_like(dtmp, delt), dtmp.abs()), -torch.minimum(torch.full_like(dtmp, delt), dtmp.abs())) / ax
        top = hsc / (cfg["cp_air"] * dpc * ax)
        kk = torch.arange(K).view(1, -1, 1, 1)
        inc = torch.where(kk < 2, top, lim)
    else:
        rdg = -cfg["rdgas"] / cfg["grav"]; k1k = cfg["akap"] / (1. - cfg["akap"]); cv_air = cfg["cp_air"] - cfg["rdgas"]
        pkz = torch.exp(k1k * torch.log(rdg * dpc / ax * ptc))
        dtmp = hsc / (cv_air * dpc)
        mn = torch.minimum(torch.full_like(dtmp, delt), dtmp.abs())
        inc = torch.where(dtmp >= 0., mn, -mn) / pkz
        kk = torch.arange(K).view(1, -1, 1, 1)
    inc = torch.where(kk < n_con, inc, torch.zeros_like(inc))
    return put(pt, *C, ptc + inc)


def dyn_core_hydro(st, g, cfg):
    """hydrostatic acoustic loop.  st: dict u v pt delp phis (pt = virtual potential temperature
    cp*..., as fv_dynamics passes it); u, v, delp, pt must already have valid halos.
    Returns dict u v pt delp mfx mfy cx cy pkz pe peln pk."""
    N = g.N
    is_, ie, js, je = 1, N, 1, N
    isd, ied, jsd, jed = g.isd, g.ied, g.jsd, g.jed
    halo, getb = halo_of(N)
    K = st["delp"].shape[1]
    n_split = cfg["n_split"]
    dt = cfg["bdt"] / n_split
    dt2 = 0.5 * dt
    cfg, cfgp = sides(cfg)
    prm = level_params(cfg, K)
    prm["hydrostatic"] = True
    pp = level_params_pert(cfgp, K, prm) if cfgp else None
    u, v, pt, delp = st["u"], st["v"], st["pt"], st["delp"]
    w = Z(u)
    hs = st["phis"]
    mfx = Z(u); mfy = Z(u); cx = Z(u); cy = Z(u)
    heat = None
    ptop, akap, cp_air = cfg["ptop"], cfg["akap"], cfg["cp_air"]
    for it in range(1, n_split + 1):
        c = c_sw(delp, pt, u, v, w, g, dt2, True, cfg["nord"])
        divgd = halo.corner(c["divg_d"]) if cfg["nord"] > 0 else c["divg_d"]
        pkc, gz, _, _, _ = geopk(c["delpc"], c["ptc"], hs, g, ptop, akap, cp_air, 1, True)
        uc, vc = p_grad_c(dt2, c["delpc"], pkc, gz, c["uc"], c["vc"], g, True)
        uc, vc = halo.cgrid(uc, vc)
        d_ext = cfg.get("d_ext", 0.0)
        dpc = a2b_ord2(delp, g) if d_ext > 0.0 else None          # delp at the corners, before d_sw (:642-644)
        d = d_sw(delp, pt, u, v, w, uc, vc, c["ua"], c["va"], divgd, g, dt, prm, pp)
        divg2 = None
        if d_ext > 0.0:           # external-mode divergence damping (:707-724): delp-weighted column mean of the corner divergence
            CC = (is_, ie + 1, js, je + 1)
            wkc = S(dpc, *CC); vtc = S(d["divg"], *CC)
            s0 = wkc[:, 0]; s1 = wkc[:, 0] * vtc[:, 0]
            for k in range(1, K):
                s0 = s0 + wkc[:, k]; s1 = s1 + wkc[:, k] * vtc[:, k]
            divg2 = put(torch.zeros_like(delp[:, :1]), *CC, ((d_ext * g.da_min_c) * s1 / s0)[:, None])
        mfx = put(mfx, is_, ie + 1, js, je, S(mfx, is_, ie + 1, js, je) + S(d["fx"], is_, ie + 1, js, je))
        mfy = put(mfy, is_, ie, js, je + 1, S(mfy, is_, ie, js, je + 1) + S(d["fy"], is_, ie, js, je + 1))
        cx = put(cx, is_, ie + 1, jsd, jed, S(cx, is_, ie + 1, jsd, jed) + S(d["crx"], is_, ie + 1, jsd, jed))
        cy = put(cy, isd, ied, js, je + 1, S(cy, isd, ied, js, je + 1) + S(d["cry"], isd, ied, js, je + 1))
        if "heat" in d:
            heat = d["heat"] if heat is None else heat + d["heat"]
        delp = halo.scalar(d["delp"]); pt = halo.scalar(d["pt"])
        pkc, gz, pe, peln, pkz = geopk(delp, pt, hs, g, ptop, akap, cp_air, 2, False)
        if cfg.get("beta", 0.0) > 0.0:      # grad1_p_update (:865-866), beta_d = 0 on the first sub-step (:373-375)
            u, v, du_dv = grad_p(d["u"], d["v"], pkc, gz, g, dt, ptop ** akap, beta=0.0 if it == 1 else cfg["beta"], du_dv=None if it == 1 else du_dv,
                                 divg2=divg2)
        else:
            u, v = grad_p(d["u"], d["v"], pkc, gz, g, dt, ptop ** akap, divg2=divg2)
        if it == n_split:
            u, v = getb(u, v)
        else:
            u, v = halo.dgrid(u, v)
    if heat is not None:
        pt = heat_update(heat, pt, delp, pkz, g, cfg, halo, True)
    return dict(u=u, v=v, pt=pt, delp=delp, mfx=mfx, mfy=mfy, cx=cx, cy=cy, pkz=pkz, pe=pe, peln=peln, pk=pkc)
