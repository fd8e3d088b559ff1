"""Driver used by make_ref_golden.py: runs the reference's DYN_CORE_TLM (model_tlmadm/dyn_core_tlm.F90:93) on all six cube tiles at once, one
Python thread per tile in lock step; the FMS halo exchanges the routine calls (start_group_halo_update, mpp_get_boundary: un-vendored FMS)
are served by the repository's own cubed-sphere index maps (oracle/cubed_sphere.py Halo, oracle/dyn_core.py GetBoundary).
Needs /root/reference: generation time only."""
import threading
import types
import numpy as np
import torch
from ref_tlm import f90py
from ref_tlm.f90py import FA

REF = "/root/reference/src/dynamics/atmos_cubed_sphere/model_tlmadm/"
NG = 3
CORNER, CGRID_NE, DGRID_NE = 101, 102, 103


class LockStep:
    """six tile threads meet at every exchange; tile 0 performs it on the six registered payloads"""

    def __init__(self, n=6):
        self.n = n
        self.barrier = threading.Barrier(n)
        self.slots = [None] * n

    def run(self, tile, payload, fn):
        self.slots[tile] = payload
        self.barrier.wait(timeout=600)
        if tile == 0:
            fn(self.slots)
        self.barrier.wait(timeout=600)


def to_oracle(fas):
    """six FA (i, j[, k]) with lower bounds (isd, jsd[, 1]) -> torch [6, K, NY, NX], NY = NX = extent of the unstaggered halo array + 1"""
    a0 = fas[0].a
    nk = a0.shape[2] if a0.ndim == 3 else 1
    n = max(a0.shape[0], a0.shape[1])
    nx = n if (a0.shape[0] != a0.shape[1] or False) else n
    return nk, nx


class Exchanger:
    def __init__(self, N):
        from oracle.dyn_core import halo_of
        self.N = N
        self.NX = N + 7
        self.halo, self.getb = halo_of(N)
        self.ls = LockStep()

    def gather(self, fas):
        a0 = fas[0].a
        nk = int(np.prod(a0.shape[2:])) if a0.ndim >= 3 else 1
        x = torch.zeros((6, nk, self.NX, self.NX), dtype=torch.float64)
        for t, f in enumerate(fas):
            a = f.a.reshape(f.a.shape[0], f.a.shape[1], -1, order="F")
            x[t, :, :a.shape[1], :a.shape[0]] = torch.from_numpy(np.ascontiguousarray(a.transpose(2, 1, 0)))
        return x

    def scatter(self, x, fas):
        for t, f in enumerate(fas):
            ni, nj = f.a.shape[0], f.a.shape[1]
            v = x[t, :, :nj, :ni].numpy().transpose(2, 1, 0)
            f.a[...] = v.reshape(f.a.shape, order="F")

    def update(self, tile, fields, kind):
        """fields: this tile's FA list, [f] or [u, v] (value and tangent lists are exchanged by separate calls)"""
        def work(slots):
            if slots[0][0].a.size == 0:
                return
            xs = [self.gather([s[i] for s in slots]) for i in range(len(slots[0]))]
            if kind == "scalar":
                ys = [self.halo.scalar(xs[0])]
            elif kind == "corner":
                ys = [self.halo.corner(xs[0])]
            elif kind == "cgrid":
                ys = list(self.halo.cgrid(xs[0], xs[1]))
            elif kind == "dgrid":
                ys = list(self.halo.dgrid(xs[0], xs[1]))
            for i, y in enumerate(ys):
                self.scatter(y, [s[i] for s in slots])
        self.ls.run(tile, fields, work)

    def update_adj(self, tile, fields, kind):
        """adjoint of update(): field_ad <- H^T field_ad, H the halo fill (torch.func.vjp of the same index maps)"""
        def work(slots):
            if slots[0][0].a.size == 0:
                return
            xs = [self.gather([s[i] for s in slots]) for i in range(len(slots[0]))]
            if kind in ("scalar", "corner"):
                fn = self.halo.scalar if kind == "scalar" else self.halo.corner
                _, vjp = torch.func.vjp(fn, torch.zeros_like(xs[0]))
                ys = [vjp(xs[0])[0]]
            else:
                fn = self.halo.cgrid if kind == "cgrid" else self.halo.dgrid
                _, vjp = torch.func.vjp(fn, torch.zeros_like(xs[0]), torch.zeros_like(xs[1]))
                ys = list(vjp((xs[0], xs[1])))
            for i, y in enumerate(ys):
                self.scatter(y, [s[i] for s in slots])
        self.ls.run(tile, fields, work)

    def boundary_adj(self, tile, u_ad, v_ad, ebuf_ad, nbuf_ad):
        """adjoint of boundary(): u_ad, v_ad += B^T (nbufferx_ad, ebuffery_ad); the buffers' adjoints are consumed"""
        N = self.N
        def fn(xu, xv):
            un, vn = self.getb(xu, xv)
            return un[:, :, N + 3, 3: N + 3], vn[:, :, 3: N + 3, N + 3]
        def work(slots):
            xu = self.gather([s[0] for s in slots]); xv = self.gather([s[1] for s in slots])
            nb = torch.stack([torch.from_numpy(np.ascontiguousarray(s[3][:N, :].T)) for s in slots])     # [6, K, N]
            eb = torch.stack([torch.from_numpy(np.ascontiguousarray(s[2][:N, :].T)) for s in slots])
            _, vjp = torch.func.vjp(fn, torch.zeros_like(xu), torch.zeros_like(xv))
            du, dv = vjp((nb, eb))
            self.scatter(xu + du, [s[0] for s in slots]); self.scatter(xv + dv, [s[1] for s in slots])
            for s in slots:
                s[2][...] = 0.0; s[3][...] = 0.0
        self.ls.run(tile, (u_ad, v_ad, ebuf_ad, nbuf_ad), work)

    def boundary(self, tile, u, v, ebuf, nbuf):
        """mpp_get_boundary(u, v, gridtype = DGRID_NE): nbufferx(1:ie-is+1, k) = u(is:ie, je+1), ebuffery(1:je-js+1, k) = v(ie+1, js:je)"""
        N = self.N
        def work(slots):
            xu = self.gather([s[0] for s in slots]); xv = self.gather([s[1] for s in slots])
            un, vn = self.getb(xu, xv)
            for t, s in enumerate(slots):
                s[3][:N, :] = un[t, :, N + 1 + 2, 1 + 2: N + 1 + 2].numpy().T        # u(is:ie, npy)
                s[2][:N, :] = vn[t, :, 1 + 2: N + 1 + 2, N + 1 + 2].numpy().T        # v(npx, js:je)
        self.ls.run(tile, (u, v, ebuf, nbuf), work)


def make_stubs(ex):
    def start_group_halo_update_tlm(pack, *args, **kw):
        domain = args[-1]
        fl = [a for a in args[:-1]]
        kind = "scalar"
        if kw.get("position") == CORNER:
            kind = "corner"
        elif kw.get("gridtype") == CGRID_NE:
            kind = "cgrid"
        elif kw.get("gridtype") == DGRID_NE:
            kind = "dgrid"
        if kind in ("cgrid", "dgrid"):
            u, u_tl, v, v_tl = fl
            ex.update(domain.tile, [u, v], kind)
            ex.update(domain.tile, [u_tl, v_tl], kind)
        else:
            f, f_tl = fl
            ex.update(domain.tile, [f], kind)
            ex.update(domain.tile, [f_tl], kind)

    def mpp_get_boundary_tlm(u, u_tl, v, v_tl, domain, ebuffery=None, ebuffery_tl=None, nbufferx=None, nbufferx_tl=None, gridtype=None):
        ex.boundary(domain.tile, u, v, ebuffery, nbufferx)
        ex.boundary(domain.tile, u_tl, v_tl, ebuffery_tl, nbufferx_tl)
    def kind_of(kw):
        return ("corner" if kw.get("position") == CORNER else "cgrid" if kw.get("gridtype") == CGRID_NE
                else "dgrid" if kw.get("gridtype") == DGRID_NE else "scalar")

    def start_group_halo_update(pack, *args, **kw):
        domain = args[-1]
        ex.update(domain.tile, list(args[:-1]), kind_of(kw))

    def start_group_halo_update_adm(pack, *args, **kw):
        domain = args[-1]
        fl = list(args[:-1])
        ex.update_adj(domain.tile, [fl[1], fl[3]] if len(fl) == 4 else [fl[1]], kind_of(kw))

    def mpp_get_boundary(u, v, domain, ebuffery=None, nbufferx=None, gridtype=None):
        ex.boundary(domain.tile, u, v, ebuffery, nbufferx)

    def mpp_get_boundary_adm(u, u_ad, v, v_ad, domain, ebuffery=None, ebuffery_ad=None, nbufferx=None, nbufferx_ad=None, gridtype=None):
        ex.boundary_adj(domain.tile, u_ad, v_ad, ebuffery_ad, nbufferx_ad)
    noop = lambda *a, **k: None
    return dict(start_group_halo_update=start_group_halo_update, start_group_halo_update_adm=start_group_halo_update_adm,
                mpp_get_boundary=mpp_get_boundary, mpp_get_boundary_adm=mpp_get_boundary_adm,
                start_group_halo_update_tlm=start_group_halo_update_tlm, complete_group_halo_update=noop, mpp_get_boundary_tlm=mpp_get_boundary_tlm,
                timing_on=noop, timing_off=noop, prt_mxm=noop, prt_maxmin=noop, is_master=lambda: False, send_data=lambda *a, **k: False,
                corner=CORNER, cgrid_ne=CGRID_NE, dgrid_ne=DGRID_NE)


def load_fill_corners(N):
    """the reference's fill_corners family (tools/fv_mp_nlm_mod.F90, model_tlmadm/fv_mp_tlm.F90: they read the module's domain indices)
    with Python dispatchers for the generic names"""
    names = ["fill_corners_2d_r8", "fill_corners_xy_2d_r8", "fill_corners_agrid_r8", "fill_corners_cgrid_r8", "fill_corners_dgrid_r8"]
    only = set(names) | {n + "_tlm" for n in names} | {n + "_adm" for n in names}
    idx = dict(is_=1, ie=N, js=1, je=N, isd=1 - NG, ied=N + NG, jsd=1 - NG, jed=N + NG, ng=NG, xdir=1, ydir=2)
    sp, fns, _ = f90py.load([REF + "../tools/fv_mp_nlm_mod.F90", REF + "fv_mp_tlm.F90", REF + "fv_mp_adm.F90"], extra=idx, only=only, defines=("SPMD",))
    g = {}
    g["fill_corners"] = lambda *a, **k: (fns["fill_corners_xy_2d_r8"] if isinstance(a[1], FA) else fns["fill_corners_2d_r8"])(*a, **k)
    g["fill_corners_tlm"] = lambda *a, **k: (fns["fill_corners_xy_2d_r8_tlm"] if isinstance(a[2], FA) else fns["fill_corners_2d_r8_tlm"])(*a, **k)
    g["fill_corners_adm"] = lambda *a, **k: (fns["fill_corners_xy_2d_r8_adm"] if isinstance(a[2], FA) else fns["fill_corners_2d_r8_adm"])(*a, **k)
    for kind in ("agrid", "cgrid", "dgrid"):
        for sfx in ("", "_tlm", "_adm"):
            if "fill_corners_%s_r8%s" % (kind, sfx) in fns:
                g["fill_corners_%s%s" % (kind, sfx)] = fns["fill_corners_%s_r8%s" % (kind, sfx)]
    g.update(xdir=1, ydir=2)
    for ns in sp.values():
        ns.update(g)
    return g


def load_reference(ex, consts, great_circle_dist, N=12):
    files = [REF + x for x in ("tp_core_tlm.F90", "sw_core_tlm.F90", "a2b_edge_tlm.F90", "dyn_core_tlm.F90", "nh_core_tlm.F90", "nh_utils_tlm.F90")]
    extra = dict(ng=NG, great_circle_dist=great_circle_dist, fpp=types.SimpleNamespace(fpp_overload_r4=False))
    extra.update(make_stubs(ex)); extra.update(consts); extra.update(load_fill_corners(N))
    return f90py.load(files, extra=extra, strict=False)


def flags_from_cfg(cfg, N):
    """fv_flags_type fields DYN_CORE_TLM reads, from the oracle's configuration dictionary (same names as the reference's namelist)"""
    d = dict(npx=N + 1, npy=N + 1, grid_type=0, k_split=1, m_split=0, d_ext=cfg.get("d_ext", 0.0), inline_q=False, fv_debug=False, a2b_ord=4, use_old_omega=False,
             use_logp=False, delt_max=1.0, d_con=cfg.get("d_con", 0.0), hydrostatic=bool(cfg.get("hydrostatic", False)), scale_z=0.0, p_fac=cfg["p_fac"],
             breed_vortex_inline=False, a_imp=cfg.get("a_imp", 1.0), nwat=0, ke_bg=0.0, fill_dp=False, do_f3d=False, convert_ke=False, beta=cfg.get("beta", 0.0),
             n_sponge=cfg.get("n_sponge", 0))
    for k in ("nord", "d2_bg", "d2_bg_k1", "d2_bg_k2", "vtdm4", "do_vort_damp", "dddmp", "d4_bg", "hord_mt", "hord_vt", "hord_tm", "hord_dp"):
        d[k] = cfg[k]
    d["hord_tr"] = cfg.get("hord_tr", cfg["hord_dp"])
    return types.SimpleNamespace(**d)


def pert_flags_from_cfg(cfg):
    """fv_flags_pert_type (model_tlmadm/fv_arrays_tlmadm.F90:37-92): the top-level keys of a two-sided configuration dictionary"""
    d = {}
    for k in ("nord", "d2_bg", "d2_bg_k1", "d2_bg_k2", "d2_bg_ks", "vtdm4", "do_vort_damp", "dddmp", "d4_bg", "hord_mt", "hord_vt", "hord_tm",
              "hord_dp", "hord_tr", "n_sponge"):
        d[k + "_pert"] = cfg[k]
    for k in ("mt", "vt", "tm", "dp", "tr"):
        d["hord_%s_ks_pert" % k] = 1
        d["hord_%s_ks_traj" % k] = 1
    d.update(split_damp=bool(cfg["split_damp"]), hord_ks_pert=bool(cfg["hord_ks_pert"]), hord_ks_traj=bool(cfg["hord_ks_traj"]))
    return types.SimpleNamespace(**d)


def run(fns, ex, grid_structs, M, N, K, f, d, cfg, ak, bk):
    """f, d: dict name -> [6, K, NY, NX] of u v w delz pt delp and phis [6, 1, NY, NX].  Returns dict of outputs [6, K(+1), NY, NX]."""
    NX = N + 7
    isd, ied, jsd, jed = 1 - NG, N + NG, 1 - NG, N + NG
    A3 = lambda nk: ((isd, ied), (jsd, jed), (1, nk))
    bnd = dict(u=((isd, ied), (jsd, jed + 1), (1, K)), v=((isd, ied + 1), (jsd, jed), (1, K)), w=A3(K), delz=A3(K), pt=A3(K), delp=A3(K),
               pe=((0, N + 1), (1, K + 1), (0, N + 1)), pk=((1, N), (1, N), (1, K + 1)), omga=A3(K), ua=A3(K), va=A3(K),
               uc=((isd, ied + 1), (jsd, jed), (1, K)), vc=((isd, ied), (jsd, jed + 1), (1, K)), mfx=((1, N + 1), (1, N), (1, K)),
               mfy=((1, N), (1, N + 1), (1, K)), cx=((1, N + 1), (jsd, jed), (1, K)), cy=((isd, ied), (1, N + 1), (1, K)),
               pkz=((1, N), (1, N), (1, K)), peln=((1, N), (1, K + 1), (1, N)), dpx=((1, N), (1, N)), ws=((1, N), (1, N)),
               gz=A3(K + 1), pkc=A3(K + 1), ptc=A3(K), crx=((1, N + 1), (jsd, jed), (1, K)), xfx=((1, N + 1), (jsd, jed), (1, K)),
               cry=((isd, ied), (1, N + 1), (1, K)), yfx=((isd, ied), (1, N + 1), (1, K)), divgd=((isd, ied + 1), (jsd, jed + 1), (1, K)),
               delpc=A3(K), ut=A3(K), vt=A3(K), zh=A3(K + 1), pk3=A3(K + 1), du=((isd, ied), (jsd, jed + 1), (1, K)),
               dv=((isd, ied + 1), (jsd, jed), (1, K)))
    order = ["u", "v", "w", "delz", "pt", "q", "delp", "pe", "pk"]
    res = [None] * 6
    errors = []

    def tile_main(t):
        try:
            bd, gs, fl0 = grid_structs(M, t, N)
            gs.area_64 = gs.area; gs.square_domain = False
            ct = dict(cfg); ct.update(cfg["traj"])
            fl = flags_from_cfg(ct, N); flp = pert_flags_from_cfg(cfg)
            a = {n: FA.alloc(b) for n, b in bnd.items()}; a_tl = {n: FA.alloc(b) for n, b in bnd.items()}
            hyd = bool(cfg.get("hydrostatic", False))
            for n in (("u", "v", "pt", "delp") if hyd else ("u", "v", "w", "delz", "pt", "delp")):
                ni, nj = a[n].a.shape[0], a[n].a.shape[1]
                a[n].a[...] = f[n][t][:, :nj, :ni].transpose(2, 1, 0)
                a_tl[n].a[...] = d[n][t][:, :nj, :ni].transpose(2, 1, 0)
            phis = FA(np.ascontiguousarray(f["phis"][t, 0][:N + 6, :N + 6].T), (isd, jsd))
            q = FA.alloc(((isd, ied), (jsd, jed), (1, K), (1, 0))); q_tl = FA.alloc(((isd, ied), (jsd, jed), (1, K), (1, 0)))
            cappa = FA.alloc(A3(K)); q_con = FA.alloc(A3(K))
            pfull = FA(np.array([0.5 * (ak[k] + ak[k + 1]) + 0.5 * (bk[k] + bk[k + 1]) * 1.e5 for k in range(K)]), (1,))
            fak = FA(np.array(ak, dtype=float), (1,)); fbk = FA(np.array(bk, dtype=float), (1,))
            i_pack = FA(np.arange(1, 13), (1,))
            domain = types.SimpleNamespace(tile=t)
            nest = types.SimpleNamespace(nest_timestep=0)
            P = lambda n: (a[n], a_tl[n])
            fns["dyn_core_tlm"](N + 1, N + 1, K, NG, 1, 0, cfg["bdt"], cfg["n_split"], 0.0, cfg["cp_air"], cfg["akap"], cappa, cfg["grav"], hyd,
                                *P("u"), *P("v"), *P("w"), *P("delz"), *P("pt"), q, q_tl, *P("delp"), *P("pe"), *P("pk"), phis, *P("ws"), *P("omga"),
                                cfg["ptop"], pfull, *P("ua"), *P("va"), *P("uc"), *P("vc"), *P("mfx"), *P("mfy"), *P("cx"), *P("cy"), *P("pkz"),
                                *P("peln"), q_con, fak, fbk, *P("dpx"), 0, gs, fl, flp, nest, types.SimpleNamespace(id_ws=0, id_zratio=0), bd, domain, True, i_pack, True,
                                *P("gz"), *P("pkc"), *P("ptc"), *P("crx"), *P("xfx"), *P("cry"), *P("yfx"), *P("divgd"), *P("delpc"), *P("ut"), *P("vt"),
                                *P("zh"), *P("pk3"), *P("du"), *P("dv"))
            res[t] = (a, a_tl)
        except BaseException as e:          # release the other tiles
            import traceback
            errors.append((t, traceback.format_exc()))
            ex.ls.barrier.abort()
    th = [threading.Thread(target=tile_main, args=(t,)) for t in range(6)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    if errors:
        raise RuntimeError("tile %d failed:\n%s" % (errors[0][0], errors[0][1]))
    out = {}
    for n in (("u", "v", "pt", "delp", "mfx", "mfy", "cx", "cy", "pkz") if cfg.get("hydrostatic", False) else
              ("u", "v", "w", "delz", "pt", "delp", "mfx", "mfy", "cx", "cy")):
        for sfx, idx in (("", 0), ("_tl", 1)):
            x = np.zeros((6, K, NX, NX))
            for t in range(6):
                fa = res[t][idx][n]
                i0, j0 = fa.lo[0], fa.lo[1]
                ni, nj = fa.a.shape[0], fa.a.shape[1]
                x[t, :, j0 + 2: j0 + 2 + nj, i0 + 2: i0 + 2 + ni] = fa.a.transpose(2, 1, 0)
            out[n + sfx] = x
    return out


def load_reference_adjoint(ex, consts, great_circle_dist, N=12):
    files = [REF + x for x in ("tp_core_adm.F90", "sw_core_adm.F90", "a2b_edge_adm.F90", "dyn_core_adm.F90", "nh_core_adm.F90", "nh_utils_adm.F90")]
    extra = dict(ng=NG, great_circle_dist=great_circle_dist, fpp=types.SimpleNamespace(fpp_overload_r4=False))
    extra.update(make_stubs(ex)); extra.update(consts); extra.update(load_fill_corners(N))
    return f90py.load(files, extra=extra, strict=False)


def run_adjoint(fns, ex, grid_structs, M, N, K, f, seed, cfg, ak, bk):
    """DYN_CORE_FWD then DYN_CORE_BWD (model_tlmadm/dyn_core_adm.F90:115, :1686) on six tiles in lock step.  seed: dict output name ->
    [6, K, NY, NX] adjoint seeds and seed["regions"]: name -> (i0, i1, j0, j1).  Returns the adjoints of the prognostic inputs."""
    NX = N + 7
    isd, ied, jsd, jed = 1 - NG, N + NG, 1 - NG, N + NG
    A3 = lambda nk: ((isd, ied), (jsd, jed), (1, nk))
    bnd = dict(u=((isd, ied), (jsd, jed + 1), (1, K)), v=((isd, ied + 1), (jsd, jed), (1, K)), w=A3(K), delz=A3(K), pt=A3(K), delp=A3(K),
               pe=((0, N + 1), (1, K + 1), (0, N + 1)), pk=((1, N), (1, N), (1, K + 1)), omga=A3(K), ua=A3(K), va=A3(K),
               uc=((isd, ied + 1), (jsd, jed), (1, K)), vc=((isd, ied), (jsd, jed + 1), (1, K)), mfx=((1, N + 1), (1, N), (1, K)),
               mfy=((1, N), (1, N + 1), (1, K)), cx=((1, N + 1), (jsd, jed), (1, K)), cy=((isd, ied), (1, N + 1), (1, K)),
               pkz=((1, N), (1, N), (1, K)), peln=((1, N), (1, K + 1), (1, N)), dpx=((1, N), (1, N)), ws=((1, N), (1, N)),
               gz=A3(K + 1), pkc=A3(K + 1), ptc=A3(K), crx=((1, N + 1), (jsd, jed), (1, K)), xfx=((1, N + 1), (jsd, jed), (1, K)),
               cry=((isd, ied), (1, N + 1), (1, K)), yfx=((isd, ied), (1, N + 1), (1, K)), divgd=((isd, ied + 1), (jsd, jed + 1), (1, K)),
               delpc=A3(K), ut=A3(K), vt=A3(K), zh=A3(K + 1), pk3=A3(K + 1), du=((isd, ied), (jsd, jed + 1), (1, K)),
               dv=((isd, ied + 1), (jsd, jed), (1, K)))
    res = [None] * 6
    errors = []
    hyd = bool(cfg.get("hydrostatic", False))
    def tile_main(t):
        try:
            bd, gs, fl0 = grid_structs(M, t, N)
            gs.area_64 = gs.area; gs.square_domain = False
            ct = dict(cfg); ct.update(cfg["traj"])
            fl = flags_from_cfg(ct, N); flp = pert_flags_from_cfg(cfg)
            a = {n: FA.alloc(b) for n, b in bnd.items()}; a_ad = {n: FA.alloc(b) for n, b in bnd.items()}
            for n in (("u", "v", "pt", "delp") if hyd else ("u", "v", "w", "delz", "pt", "delp")):
                ni, nj = a[n].a.shape[0], a[n].a.shape[1]
                a[n].a[...] = f[n][t][:, :nj, :ni].transpose(2, 1, 0)
            phis = FA(np.ascontiguousarray(f["phis"][t, 0][:N + 6, :N + 6].T), (isd, jsd))
            q = FA.alloc(((isd, ied), (jsd, jed), (1, K), (1, 0))); q_ad = FA.alloc(((isd, ied), (jsd, jed), (1, K), (1, 0)))
            cappa = FA.alloc(A3(K)); q_con = FA.alloc(A3(K))
            pfull = FA(np.array([0.5 * (ak[k] + ak[k + 1]) + 0.5 * (bk[k] + bk[k + 1]) * 1.e5 for k in range(K)]), (1,))
            fak = FA(np.array(ak, dtype=float), (1,)); fbk = FA(np.array(bk, dtype=float), (1,))
            i_pack = FA(np.arange(1, 13), (1,))
            domain = types.SimpleNamespace(tile=t)
            nest = types.SimpleNamespace(nest_timestep=0)
            idiag = types.SimpleNamespace(id_ws=0, id_zratio=0)
            V = lambda n: a[n]
            P = lambda n: (a[n], a_ad[n])
            assert not f90py.stack()
            fns["dyn_core_fwd"](N + 1, N + 1, K, NG, 1, 0, cfg["bdt"], cfg["n_split"], 0.0, cfg["cp_air"], cfg["akap"], cappa, cfg["grav"], hyd,
                                V("u"), V("v"), V("w"), V("delz"), V("pt"), q, V("delp"), V("pe"), V("pk"), phis, V("ws"), V("omga"),
                                cfg["ptop"], pfull, V("ua"), V("va"), V("uc"), V("vc"), V("mfx"), V("mfy"), V("cx"), V("cy"), V("pkz"),
                                V("peln"), q_con, fak, fbk, V("dpx"), 0, gs, fl, flp, nest, idiag, bd, domain, True, i_pack, True,
                                V("gz"), V("pkc"), V("ptc"), V("crx"), V("xfx"), V("cry"), V("yfx"), V("divgd"), V("delpc"), V("ut"), V("vt"),
                                V("zh"), V("pk3"), V("du"), V("dv"))
            for n, rg in seed["regions"].items():
                i0, i1, j0, j1 = rg
                fa_ = a_ad[n]
                li, lj = fa_.lo[0], fa_.lo[1]
                # seed[n][t]: [K, NY, NX] oracle layout (index = coordinate + 2)
                fa_.a[i0 - li: i1 - li + 1, j0 - lj: j1 - lj + 1, :] = seed[n][t][:, j0 + 2: j1 + 3, i0 + 2: i1 + 3].transpose(2, 1, 0)
            fns["dyn_core_bwd"](N + 1, N + 1, K, NG, 1, 0, cfg["bdt"], cfg["n_split"], 0.0, cfg["cp_air"], cfg["akap"], cappa, cfg["grav"], hyd,
                                *P("u"), *P("v"), *P("w"), *P("delz"), *P("pt"), q, q_ad, *P("delp"), *P("pe"), *P("pk"), phis, *P("ws"), *P("omga"),
                                cfg["ptop"], pfull, *P("ua"), *P("va"), *P("uc"), *P("vc"), *P("mfx"), *P("mfy"), *P("cx"), *P("cy"), *P("pkz"),
                                *P("peln"), q_con, fak, fbk, *P("dpx"), 0, gs, fl, flp, nest, idiag, bd, domain, True, i_pack, True,
                                *P("gz"), *P("pkc"), *P("ptc"), *P("crx"), *P("xfx"), *P("cry"), *P("yfx"), *P("divgd"), *P("delpc"), *P("ut"), *P("vt"),
                                *P("zh"), *P("pk3"), *P("du"), *P("dv"))
            assert not f90py.stack(), "checkpoint stack not empty after the backward sweep"
            res[t] = (a, a_ad)
        except BaseException as e:
            import traceback
            errors.append((t, traceback.format_exc()))
            ex.ls.barrier.abort()
    th = [threading.Thread(target=tile_main, args=(t,)) for t in range(6)]
    for x in th: x.start()
    for x in th: x.join()
    if errors:
        raise RuntimeError("tile %d failed:\n%s" % (errors[0][0], errors[0][1]))
    out = {}
    for n in (("u", "v", "pt", "delp") if hyd else ("u", "v", "w", "delz", "pt", "delp")):
        x = np.zeros((6, K, NX, NX))
        for t in range(6):
            fa_ = res[t][1][n]
            i0, j0 = fa_.lo[0], fa_.lo[1]
            ni, nj = fa_.a.shape[0], fa_.a.shape[1]
            x[t, :, j0 + 2: j0 + 2 + nj, i0 + 2: i0 + 2 + ni] = fa_.a.transpose(2, 1, 0)
        out[n + "_ad"] = x
    return out
