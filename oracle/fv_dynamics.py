"""ORACLE (test infrastructure only).  fv_dynamics (model step driver), tracer_2d and the
fv3jedi_lm dynamics-component step, restated in torch float64 from

  model/fv_dynamics_nlm.F90 :70-760  (TL model_tlmadm/fv_dynamics_tlm.F90:87, AD fv_dynamics_adm.F90:110/874)
  model/fv_tracer2d_nlm.F90 :275-516 tracer_2d (q_split = 1: one sub-step; q_split = 0: sub-steps from the Courant numbers)
  src/dynamics/fv3jedi_lm_dynamics_mod.F90  step_nl/tl/ad :268/:347/:460, traj_to_fv3 :717,
                                            pert_to_fv3 :848, fv3_to_pert :893
  model_tlmadm/fv_pressure.F90 :22-69 compute_fv3_pressures

Configuration: adiabatic = F, consv_te = 0, tau = 0, consv_am = F, no omega diagnostics,
remap_option = 0, |kord| = 17, moist_phys = T, nwat = 3 (sphum = tracer 1).

The TL oracle is torch.func.jvp of step_nl with respect to the prognostic inputs and the AD
oracle torch.func.vjp -- i.e. exactly TL = dN/dx and AD = (dN/dx)^T for split_* = .false.
(SURVEY fact 5).  parity unpinned (no reference vectors).
"""
import numpy as np
import torch
from .cubed_sphere import R
from .sw_core import S, Z
from .sw_core import put as put4
from .fv_mapz import put
from . import tp_core as tp
from .dyn_core import halo_of, dyn_core_hydro
from .fv_mapz import lagrangian_to_eulerian


def tracer_2d(qs, dp1, mfx, mfy, cx, cy, g, hord, hord_pert=None, q_split=1, q_split_max=3, halo=None):
    """fv_tracer2d_nlm.F90:275-516.  qs: list of tracers (halo valid).
    q_split = 1: one sub-step (nsplt = 1).
    q_split = 0: nsplt = int(1 + max_k cmax(k)) and ksplt(k) = int(1 + cmax(k)) from the Courant numbers (:351-378), the
    transport terms scaled by 1/ksplt(k) (:384-420), level k advanced only in sub-steps it <= ksplt(k) (:433), dp1 carried forward
    except after the last sub-step (:482-488), halo of q renewed between sub-steps (:496-502, `halo`).  ksplt / frac are not
    differentiated (fv_tracer2d_tlm.F90:964-996 scales cx_tl ... by the same frac).  The loop here always runs q_split_max
    sub-steps with a level mask (sub-steps beyond nsplt change nothing), so the function stays traceable by torch.func."""
    N = g.N
    is_, ie, js, je = 1, N, 1, N
    isd, ied, jsd, jed = g.isd, g.ied, g.jsd, g.jed
    sin1 = g.sin_sg[..., 1]; sin2 = g.sin_sg[..., 2]; sin3 = g.sin_sg[..., 3]; sin4 = g.sin_sg[..., 4]
    i0, i1, j0, j1 = is_, ie + 1, jsd, jed
    c = S(cx, i0, i1, j0, j1)
    xfx = put4(Z(cx), i0, i1, j0, j1, torch.where(c > 0., c * S(g.dxa, i0 - 1, i1 - 1, j0, j1) * S(g.dy, i0, i1, j0, j1) * S(sin3, i0 - 1, i1 - 1, j0, j1),
                                                 c * S(g.dxa, i0, i1, j0, j1) * S(g.dy, i0, i1, j0, j1) * S(sin1, i0, i1, j0, j1)))
    i0, i1, j0, j1 = isd, ied, js, je + 1
    c = S(cy, i0, i1, j0, j1)
    yfx = put4(Z(cx), i0, i1, j0, j1, torch.where(c > 0., c * S(g.dya, i0, i1, j0 - 1, j1 - 1) * S(g.dx, i0, i1, j0, j1) * S(sin4, i0, i1, j0 - 1, j1 - 1),
                                                 c * S(g.dya, i0, i1, j0, j1) * S(g.dx, i0, i1, j0, j1) * S(sin2, i0, i1, j0, j1)))
    rar = S(g.rarea, is_, ie, js, je)
    nit = 1
    ksplt = nsplt = None
    if q_split == 0:
        K = cx.shape[1]
        cm = torch.maximum(S(cx, is_, ie, js, je).detach().abs(), S(cy, is_, ie, js, je).detach().abs())
        lower = (torch.arange(1, K + 1) >= K // 6).to(cm.dtype)[None, :, None, None]          # k < npz/6: plain maximum (:353-365)
        cm = cm + lower * (1. - S(g.sin_sg[..., 5], is_, ie, js, je))
        cmax = cm.amax(dim=(0, 2, 3))                                                          # all tiles = mp_reduce_max (:374)
        nsplt = torch.floor(1. + cmax.max())
        ksplt = torch.where(nsplt != 1., torch.floor(1. + cmax), torch.ones_like(cmax))[None, :, None, None]
        frac = 1. / ksplt
        cx, cy, xfx, yfx, mfx, mfy = cx * frac, cy * frac, xfx * frac, yfx * frac, mfx * frac, mfy * frac
        nit = q_split_max
    ra_x = put4(Z(cx), is_, ie, jsd, jed, S(g.area, is_, ie, jsd, jed) + (S(xfx, is_, ie, jsd, jed) - S(xfx, is_ + 1, ie + 1, jsd, jed)))
    ra_y = put4(Z(cx), isd, ied, js, je, S(g.area, isd, ied, js, je) + (S(yfx, isd, ied, js, je) - S(yfx, isd, ied, js + 1, je + 1)))
    qs = list(qs)
    for it in range(1, nit + 1):
        if it > 1:
            qs = [halo.scalar(q) for q in qs]
        dp2 = S(dp1, is_, ie, js, je) + ((S(mfx, is_, ie, js, je) - S(mfx, is_ + 1, ie + 1, js, je)) + (S(mfy, is_, ie, js, je) - S(mfy, is_, ie, js + 1, je + 1))) * rar
        out = []
        for q in qs:
            fx, fy, _ = tp.fv_tp_2d(q, cx, cy, hord, xfx, yfx, g, ra_x, ra_y, mfx=mfx, mfy=mfy)
            if hord_pert is not None and hord_pert != hord:      # model_tlmadm/fv_tracer2d_tlm.F90:1060-1090
                from .d_sw import splice
                fxp, fyp, _ = tp.fv_tp_2d(q, cx, cy, hord_pert, xfx, yfx, g, ra_x, ra_y, mfx=mfx, mfy=mfy)
                fx, fy = splice(fxp, fx), splice(fyp, fy)
            qn = (S(q, is_, ie, js, je) * S(dp1, is_, ie, js, je) +
                  ((S(fx, is_, ie, js, je) - S(fx, is_ + 1, ie + 1, js, je)) + (S(fy, is_, ie, js, je) - S(fy, is_, ie, js + 1, je + 1))) * rar) / dp2
            if ksplt is not None:
                qn = torch.where(it <= ksplt, qn, S(q, is_, ie, js, je))
            out.append(put4(q, is_, ie, js, je, qn))
        qs = out
        if ksplt is not None:
            take = (it <= ksplt) & (nsplt != it)
            dp1 = put4(dp1, is_, ie, js, je, torch.where(take, dp2, S(dp1, is_, ie, js, je)))
    if q_split == 0:
        tracer_2d.last_nsplt = int(nsplt)            # (tests: which branch the inputs exercised, and the overflow condition)
    return qs


def compute_pressures(delp, ptop, kappa):
    """fv_pressure.F90:22-69 (compute domain values; arrays are full size)"""
    z1 = torch.zeros_like(delp[:, :1])
    pe = torch.cumsum(torch.cat([z1 + ptop, delp], dim=1), dim=1)
    peln = torch.log(pe)
    pk = torch.exp(kappa * peln)
    pkz = (pk[:, 1:] - pk[:, :-1]) / (kappa * (peln[:, 1:] - peln[:, :-1]))
    return pe, pk, pkz, peln


def fv_dynamics(st, g, ak, bk, cfg):
    """st: u v pt(=T) delp q(list) phis pkz pk pe peln [w delz].  Returns the same keys."""
    N = g.N
    is_, ie, js, je = 1, N, 1, N
    halo, getb = halo_of(N)
    zvir, kappa = cfg["zvir"], cfg["akap"]
    hydro = cfg["hydrostatic"]
    k_split = cfg["k_split"]
    C = (slice(None), slice(None), R(js, je), R(is_, ie))
    u, v, pt, delp, q = st["u"], st["v"], st["pt"], st["delp"], list(st["q"])
    pkz, pk, pe, peln = st["pkz"], st["pk"], st["pe"], st["peln"]
    dp1c = zvir * q[0][C]
    if not hydro:
        rdg = -cfg["rdgas"] / cfg["grav"]
        pkz = put(pkz, C, torch.exp(kappa * torch.log(rdg * delp[C] * pt[C] * (1. + dp1c) / st["delz"][C])))
    pt = put(pt, C, pt[C] * (1. + dp1c) / pkz[C])
    cfgd = dict(cfg); cfgd["bdt"] = cfg["dt"] / k_split
    w = st.get("w"); delz = st.get("delz")
    for n_map in range(1, k_split + 1):
        delp = halo.scalar(delp); pt = halo.scalar(pt)
        u, v = halo.dgrid(u, v)
        dp1 = delp
        last_step = (n_map == k_split)
        if hydro:
            d = dyn_core_hydro(dict(u=u, v=v, pt=pt, delp=delp, phis=st["phis"]), g, cfgd)
        else:
            from .nh import dyn_core_nh
            d = dyn_core_nh(dict(u=u, v=v, pt=pt, delp=delp, w=w, delz=delz, phis=st["phis"]), g, cfgd, ak, bk, n_map == 1)
            w, delz = d["w"], d["delz"]
        u, v, pt, delp = d["u"], d["v"], d["pt"], d["delp"]
        q = [halo.scalar(x) for x in q]
        two = bool(cfg.get("traj"))
        q = tracer_2d(q, dp1, d["mfx"], d["mfy"], d["cx"], d["cy"], g, cfg["traj"]["hord_tr"] if two else cfg["hord_tr"],
                      cfg["hord_tr"] if two else None, q_split=cfg.get("q_split", 1), q_split_max=cfg.get("q_split_max", 3), halo=halo)
        r = dict(pe=d["pe"], pk=d["pk"], peln=d["peln"], pkz=d["pkz"], delp=delp, pt=pt, u=u, v=v, q=q)
        if not hydro:
            r.update(w=w, delz=delz, ws=d["ws"])
        r = lagrangian_to_eulerian(r, g, ak, bk, cfg, last_step)
        u, v, pt, delp, q = r["u"], r["v"], r["pt"], r["delp"], r["q"]
        pe, pk, peln, pkz = r["pe"], r["pk"], r["peln"], r["pkz"]
        if not hydro:
            w, delz = r["w"], r["delz"]
    out = dict(u=u, v=v, pt=pt, delp=delp, q=q, pe=pe, pk=pk, peln=peln, pkz=pkz)
    if not hydro:
        out.update(w=w, delz=delz)
    return out


def step_nl(x, g, ak, bk, cfg, phis, winds=False):
    """fv3jedi_lm dynamics step (fv3jedi_lm_dynamics_mod.F90:268-345): x = dict of compute-domain
    prognostics u v t delp qv ql qi o3 [w delz] stored in full-size arrays (halo ignored).
    Returns the same dict after one model step (winds: + the A-grid lon / lat winds ua, va, :839-840)."""
    N = g.N
    halo, getb = halo_of(N)
    hydro = cfg["hydrostatic"]
    C = (slice(None), slice(None), R(1, N), R(1, N))
    def dom(a):
        return put(torch.zeros_like(a), C, a[C])
    u, v = dom(x["u"]), dom(x["v"])
    u, v = getb(u, v)
    delp = dom(x["delp"])
    pe, pk, pkz, peln = compute_pressures(delp, cfg["ptop"], cfg["akap"])
    pe, pk, pkz, peln = dom(pe), dom(pk), dom(pkz), dom(peln)
    st = dict(u=u, v=v, pt=dom(x["t"]), delp=delp, q=[dom(x[n]) for n in ("qv", "ql", "qi", "o3")],
              phis=halo.scalar(phis), pe=pe, pk=pk, pkz=pkz, peln=peln)
    if not hydro:
        st["w"] = dom(x["w"]); st["delz"] = dom(x["delz"])
    o = fv_dynamics(st, g, ak, bk, cfg)
    out = dict(u=dom(o["u"]), v=dom(o["v"]), t=dom(o["pt"]), delp=dom(o["delp"]))
    for n, qq in zip(("qv", "ql", "qi", "o3"), o["q"]):
        out[n] = dom(qq)
    if not hydro:
        out["w"] = dom(o["w"]); out["delz"] = dom(o["delz"])
    if winds:       # cubed_to_latlon(..., mode = 1, c2l_ord = 4) at the end of fv_dynamics (fv_dynamics_nlm.F90:738) -> traj%ua, va
        from .c2l import c2l_ord4
        uh, vh = halo.dgrid(o["u"], o["v"])
        out["ua"], out["va"] = c2l_ord4(uh, vh, g)
    return out
