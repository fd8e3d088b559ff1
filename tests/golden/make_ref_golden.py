"""Generates tests/golden/ref_*.npz: outputs of the REFERENCE'S OWN Fortran routines, executed IN THIS CONTAINER from /root/reference by the
Fortran -> Python transpiler tests/ref_tlm/f90py.py (no Fortran compiler, FMS or MPI exists in the image, so the reference cannot be
compiled; the transpiler runs its source statement by statement in float64).  Run once here:

    python tests/golden/make_ref_golden.py            # writes tests/golden/ref_d_sw_tlm.npz, ...

The fixtures hold the seeded inputs and every output (values and tangents); tests/test_ref_golden.py compares the oracle (torch.func.jvp of
the restated primal) and the library (host emulation through the C ABI) with them.  /root/reference is needed only by this script.
The grid metrics are the repository's own (oracle/grid.py): the reference reads them from FMS at run time."""
import os
import sys
import types
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "fv3-jedi-linearmodel_b200"))
from ref_tlm import f90py                      # noqa: E402
from ref_tlm.f90py import FA                   # noqa: E402

REF = "/root/reference/src/dynamics/atmos_cubed_sphere/model_tlmadm/"
NG = 3


def great_circle_dist(q1, q2, radius=None):
    """fv_grid_utils_nlm great_circle_dist (haversine); q1, q2: (lon, lat) sections"""
    p1 = np.asarray(q1.a if isinstance(q1, FA) else q1); p2 = np.asarray(q2.a if isinstance(q2, FA) else q2)
    beta = np.arcsin(np.sqrt(np.sin((p1[1] - p2[1]) / 2.) ** 2 + np.cos(p1[1]) * np.cos(p2[1]) * np.sin((p1[0] - p2[0]) / 2.) ** 2)) * 2.
    return beta if radius is None else radius * beta


def load_reference():
    only = {"xppm_tlm", "yppm_tlm", "fv_tp_2d_tlm", "copy_corners_tlm", "deln_flux_tlm", "d_sw_tlm", "xtp_u_tlm", "ytp_v_tlm", "del6_vt_flux_tlm",
            "compute_divergence_damping_tlm", "smag_corner_tlm", "fv_tp_2d", "xppm", "yppm", "copy_corners", "deln_flux", "pert_ppm", "xtp_u", "ytp_v",
            "del6_vt_flux", "compute_divergence_damping", "smag_corner", "a2b_ord4_tlm", "a2b_ord4", "extrap_corner_tlm", "extrap_corner",
            "fill_corners_tlm", "fill_corners", "c_sw_tlm", "d2a2c_vect_tlm", "divergence_corner_tlm", "fill_4corners_tlm", "fill2_4corners_tlm",
            "edge_interpolate4_tlm", "edge_interpolate4"}
    extra = dict(ng=NG, great_circle_dist=great_circle_dist, fpp=types.SimpleNamespace(fpp_overload_r4=False))
    return f90py.load([REF + "tp_core_tlm.F90", REF + "sw_core_tlm.F90", REF + "a2b_edge_tlm.F90"], extra=extra, only=only)


def fa(a, i0, i1, j0, j1):
    """[NY, NX(, c)] array with Fortran (i, j) at [j + 2, i + 2] -> FA over (i0:i1, j0:j1(, 1:c))"""
    s = a[j0 + 2: j1 + 3, i0 + 2: i1 + 3]
    if s.ndim == 3:
        return FA(np.ascontiguousarray(s.transpose(1, 0, 2)), (i0, j0, 1))
    return FA(np.ascontiguousarray(s.T), (i0, j0))


def back(f, N):
    """FA over any (i0:i1, j0:j1) -> [NY, NX] array, zero outside"""
    out = np.zeros((N + 7, N + 7))
    i0, j0 = f.lo
    ni, nj = f.a.shape
    out[j0 + 2: j0 + 2 + nj, i0 + 2: i0 + 2 + ni] = f.a.T
    return out


def grid_structs(M, t, N):
    bd = types.SimpleNamespace(is_=1, ie=N, js=1, je=N, isd=1 - NG, ied=N + NG, jsd=1 - NG, jed=N + NG, ng=NG)
    gs = types.SimpleNamespace(nested=False, grid_type=0, stretched_grid=False, sw_corner=True, se_corner=True, nw_corner=True, ne_corner=True,
                               da_min=M["da_min"], da_min_c=M["da_min_c"])
    isd, ied, jsd, jed = bd.isd, bd.ied, bd.jsd, bd.jed
    A = (isd, ied, jsd, jed); B = (isd, ied + 1, jsd, jed + 1); U = (isd, ied, jsd, jed + 1); V = (isd, ied + 1, jsd, jed)
    for k, b in dict(area=A, rarea=A, dxa=A, dya=A, rdxa=A, rdya=A, cosa_s=A, rsin2=A, dx=U, rdx=U, dy=V, rdy=V, dxc=V, rdxc=V, dyc=U, rdyc=U,
                     area_c=B, rarea_c=B, cosa=B, sina=B, rsina=(1, N + 1, 1, N + 1), cosa_u=V, sina_u=V, rsin_u=V, cosa_v=U, sina_v=U, rsin_v=U,
                     f0=A, fC=B, divg_u=U, divg_v=V, del6_u=U, del6_v=V).items():
        setattr(gs, k.lower(), fa(M[k][t], *b))
    gs.fc = gs.fc if hasattr(gs, "fc") else None
    gs.sin_sg = FA(np.ascontiguousarray(M["sin_sg"][t, :N + 6, :N + 6, 1:].transpose(1, 0, 2)), (isd, jsd, 1))
    gs.cos_sg = FA(np.ascontiguousarray(M["cos_sg"][t, :N + 6, :N + 6, 1:].transpose(1, 0, 2)), (isd, jsd, 1))
    gs.grid = fa(M["grid"][t], *B); gs.agrid = fa(M["agrid"][t], *A)
    for e in ("edge_w", "edge_e", "edge_s", "edge_n"):
        setattr(gs, e, FA(np.ascontiguousarray(M[e][t, 3:3 + N + 1]), (1,)))     # Fortran index j at [j + 2]
    fl = types.SimpleNamespace(npx=N + 1, npy=N + 1, grid_type=0, do_f3d=False)
    return bd, gs, fl


D_SW_BOUNDS = None


def d_sw_case(fns, M, N, t, f, d, prm, dt, hydrostatic, d_con=0.0):
    """one level of D_SW_TLM on tile t.  f, d: dict name -> [NY, NX] value / tangent of delp pt u v w uc vc ua va divg_d"""
    bd, gs, fl = grid_structs(M, t, N)
    isd, ied, jsd, jed = bd.isd, bd.ied, bd.jsd, bd.jed
    A = (isd, ied, jsd, jed)
    bnd = dict(delp=A, pt=A, w=A, ua=A, va=A, u=(isd, ied, jsd, jed + 1), vc=(isd, ied, jsd, jed + 1), v=(isd, ied + 1, jsd, jed),
               uc=(isd, ied + 1, jsd, jed), divg_d=(isd, ied + 1, jsd, jed + 1))
    a = {n: fa(f[n], *bnd[n]) for n in bnd}; a_tl = {n: fa(d[n], *bnd[n]) for n in bnd}
    Z = lambda b: FA.alloc(((b[0], b[1]), (b[2], b[3])))
    X = (1, N + 1, jsd, jed); Y = (isd, ied, 1, N + 1); C = (1, N, 1, N)
    o = {}
    for n, b in dict(delpc=A, ptc=A, xflux=(1, N + 1, 1, N), yflux=(1, N, 1, N + 1), cx=X, cy=Y, crx_adv=X, cry_adv=Y, xfx_adv=X, yfx_adv=Y,
                     heat_source=C, dpx=C, z_rat=A, q_con=A).items():
        o[n] = Z(b); o[n + "_tl"] = Z(b)
    q = FA.alloc(((isd, ied), (jsd, jed), (1, 1), (1, 1))); q_tl = FA.alloc(((isd, ied), (jsd, jed), (1, 1), (1, 1)))
    p = prm
    fns["d_sw_tlm"](o["delpc"], o["delpc_tl"], a["delp"], a_tl["delp"], o["ptc"], o["ptc_tl"], a["pt"], a_tl["pt"], a["u"], a_tl["u"], a["v"], a_tl["v"],
                    a["w"], a_tl["w"], a["uc"], a_tl["uc"], a["vc"], a_tl["vc"], a["ua"], a_tl["ua"], a["va"], a_tl["va"], a["divg_d"], a_tl["divg_d"],
                    o["xflux"], o["xflux_tl"], o["yflux"], o["yflux_tl"], o["cx"], o["cx_tl"], o["cy"], o["cy_tl"], o["crx_adv"], o["crx_adv_tl"],
                    o["cry_adv"], o["cry_adv_tl"], o["xfx_adv"], o["xfx_adv_tl"], o["yfx_adv"], o["yfx_adv_tl"], o["q_con"], o["z_rat"], o["z_rat_tl"],
                    0.0, o["heat_source"], o["heat_source_tl"], o["dpx"], o["dpx_tl"], 0.0, 1, 1, q, q_tl, 1, 1, False, dt,
                    p["hord_tr"], p["hord_mt"], p["hord_vt"], p["hord_tm"], p["hord_dp"], p["nord"], p["nord_v"], p["nord_w"], p["nord_t"],
                    p["dddmp"], p["d2_bg"], p["d4_bg"], p["damp_v"], p["damp_w"], p["damp_t"], d_con, hydrostatic, gs, fl, bd,
                    p["hord_tr"], p["hord_mt"], p["hord_vt"], p["hord_tm"], p["hord_dp"], False, p["nord"], p["nord_v"], p["nord_w"], p["nord_t"],
                    p["dddmp"], p["d2_bg"], p["d4_bg"], p["damp_v"], p["damp_w"], p["damp_t"])
    res = {}
    for n in ("delp", "pt", "u", "v", "w"):
        res[n] = back(a[n], N); res[n + "_tl"] = back(a_tl[n], N)
    for n, key in (("xflux", "fx"), ("yflux", "fy"), ("crx_adv", "crx"), ("cry_adv", "cry"), ("xfx_adv", "xfx"), ("yfx_adv", "yfx"), ("heat_source", "heat")):
        res[key] = back(o[n], N); res[key + "_tl"] = back(o[n + "_tl"], N)
    return res


def load_reference_adjoint():
    extra = dict(ng=NG, great_circle_dist=great_circle_dist, fpp=types.SimpleNamespace(fpp_overload_r4=False))
    return f90py.load([REF + "tp_core_adm.F90", REF + "sw_core_adm.F90", REF + "a2b_edge_adm.F90"], extra=extra, strict=False)


def d_sw_adjoint_case(fns, M, N, t, f, seed, prm, dt, hydrostatic, d_con=0.0):
    """D_SW_FWD then D_SW_BWD (model_tlmadm/sw_core_adm.F90:1773, :3269) on tile t.  seed: dict output name -> [NY, NX] adjoint seed of
    delp pt u v w (in-place outputs) and fx fy crx cry xfx yfx heat.  Returns the adjoints of the ten inputs."""
    bd, gs, fl = grid_structs(M, t, N)
    isd, ied, jsd, jed = bd.isd, bd.ied, bd.jsd, bd.jed
    A = (isd, ied, jsd, jed)
    bnd = dict(delp=A, pt=A, w=A, ua=A, va=A, u=(isd, ied, jsd, jed + 1), vc=(isd, ied, jsd, jed + 1), v=(isd, ied + 1, jsd, jed),
               uc=(isd, ied + 1, jsd, jed), divg_d=(isd, ied + 1, jsd, jed + 1))
    a = {n: fa(f[n], *bnd[n]) for n in bnd}
    Z = lambda b: FA.alloc(((b[0], b[1]), (b[2], b[3])))
    X = (1, N + 1, jsd, jed); Y = (isd, ied, 1, N + 1); C = (1, N, 1, N)
    ob = dict(delpc=A, ptc=A, xflux=(1, N + 1, 1, N), yflux=(1, N, 1, N + 1), cx=X, cy=Y, crx_adv=X, cry_adv=Y, xfx_adv=X, yfx_adv=Y,
              heat_source=C, dpx=C, z_rat=A, q_con=A)
    o = {n: Z(b) for n, b in ob.items()}
    q = FA.alloc(((isd, ied), (jsd, jed), (1, 1), (1, 1))); q_ad = FA.alloc(((isd, ied), (jsd, jed), (1, 1), (1, 1)))
    p = prm
    sw = (p["hord_tr"], p["hord_mt"], p["hord_vt"], p["hord_tm"], p["hord_dp"], p["nord"], p["nord_v"], p["nord_w"], p["nord_t"],
          p["dddmp"], p["d2_bg"], p["d4_bg"], p["damp_v"], p["damp_w"], p["damp_t"], d_con, hydrostatic, gs, fl, bd,
          p["hord_tr"], p["hord_mt"], p["hord_vt"], p["hord_tm"], p["hord_dp"], False, p["nord"], p["nord_v"], p["nord_w"], p["nord_t"],
          p["dddmp"], p["d2_bg"], p["d4_bg"], p["damp_v"], p["damp_w"], p["damp_t"])
    stack = f90py.stack()
    assert not stack
    fns["d_sw_fwd"](o["delpc"], a["delp"], o["ptc"], a["pt"], a["u"], a["v"], a["w"], a["uc"], a["vc"], a["ua"], a["va"], a["divg_d"],
                    o["xflux"], o["yflux"], o["cx"], o["cy"], o["crx_adv"], o["cry_adv"], o["xfx_adv"], o["yfx_adv"], o["q_con"], o["z_rat"],
                    0.0, o["heat_source"], o["dpx"], 0.0, 1, 1, q, 1, 1, False, dt, *sw)
    ad = {n: Z(b) for n, b in bnd.items()}
    oad = {n: Z(b) for n, b in ob.items()}
    for n in ("delp", "pt", "u", "v", "w"):
        ad[n] = fa(seed[n], *bnd[n])
    for n, key in (("xflux", "fx"), ("yflux", "fy"), ("crx_adv", "crx"), ("cry_adv", "cry"), ("xfx_adv", "xfx"), ("yfx_adv", "yfx"), ("heat_source", "heat")):
        oad[n] = fa(seed[key], *ob[n])
    fns["d_sw_bwd"](o["delpc"], oad["delpc"], a["delp"], ad["delp"], o["ptc"], oad["ptc"], a["pt"], ad["pt"], a["u"], ad["u"], a["v"], ad["v"],
                    a["w"], ad["w"], a["uc"], ad["uc"], a["vc"], ad["vc"], a["ua"], ad["ua"], a["va"], ad["va"], a["divg_d"], ad["divg_d"],
                    o["xflux"], oad["xflux"], o["yflux"], oad["yflux"], o["cx"], oad["cx"], o["cy"], oad["cy"], o["crx_adv"], oad["crx_adv"],
                    o["cry_adv"], oad["cry_adv"], o["xfx_adv"], oad["xfx_adv"], o["yfx_adv"], oad["yfx_adv"], o["q_con"], o["z_rat"], oad["z_rat"],
                    0.0, o["heat_source"], oad["heat_source"], o["dpx"], oad["dpx"], 0.0, 1, 1, q, q_ad, 1, 1, False, dt, *sw)
    assert not stack, ("checkpoint stack not empty after the backward sweep", len(stack))
    return {n + "_ad": back(ad[n], N) for n in bnd}


def main():
    from common import metrics, rnd
    from test_d_sw import dsw_inputs
    spaces, fns, src = load_reference()
    N, K = 12, 1
    M = metrics(N)
    f, rng = dsw_inputs(N, K, 11)
    d = {n: 1e-2 * np.abs(f[n]).std() * rnd(rng, N, K) for n in f}
    base = dict(hord_tr=2, hord_mt=2, hord_vt=2, hord_tm=2, hord_dp=2, nord=1, nord_v=1, nord_w=1, nord_t=1, d2_bg=0.015, damp_v=0.0005,
                damp_w=0.0005, damp_t=0.0005, dddmp=0.2, d4_bg=0.15)
    sponge = dict(base, hord_mt=1, hord_vt=1, hord_tm=1, hord_dp=1, nord=0, d2_bg=4.0, nord_w=0, damp_w=4.0, nord_v=0, damp_v=2.0, nord_t=0)
    cases = dict(nonhydro=(base, False, 0.0), sponge=(sponge, False, 0.0), hydro=(base, True, 0.0), heating=(base, False, 0.5),
                 ord333=(dict(base, hord_mt=333, hord_vt=333, hord_tm=333, hord_dp=333), False, 0.0))
    out = {"N": N, "dt": 450.0}
    for n in f:
        out["in." + n] = f[n]; out["in_tl." + n] = d[n]
    tiles = (0, 3)
    out["tiles"] = np.array(tiles)
    for name, (prm, hydro, d_con) in cases.items():
        for t in tiles:
            r = d_sw_case(fns, M, N, t, {n: f[n][t, 0] for n in f}, {n: d[n][t, 0] for n in f}, prm, 450.0, hydro, d_con)
            for k, v in r.items():
                out["%s.t%d.%s" % (name, t, k)] = v
        print("d_sw_tlm", name, "done")
    # adjoint: the reference's own reverse-mode code (D_SW_FWD / D_SW_BWD with Tapenade's checkpoint stack)
    sp_a, fns_a, src_a = load_reference_adjoint()
    from common import region
    npx = N + 1
    C = (1, N, 1, N)
    regs = dict(delp=C, pt=C, w=C, heat=C, u=(1, N, 1, npx), v=(1, npx, 1, N), fx=(1, npx, 1, N), fy=(1, N, 1, npx),
                crx=(1, npx, -2, N + 3), xfx=(1, npx, -2, N + 3), cry=(-2, N + 3, 1, npx), yfx=(-2, N + 3, 1, npx))
    seed = {}
    for n, rg in regs.items():
        y = np.zeros((6, 1, N + 7, N + 7))
        region(y, *rg)[...] = region(rnd(rng, N, 1), *rg)
        seed[n] = y
        out["seed." + n] = y
    for name, (prm, hydro, d_con) in cases.items():
        sd = {n: (v if not ((hydro and n == "w") or (d_con == 0.0 and n == "heat")) else 0 * v) for n, v in seed.items()}
        for t in tiles:
            r = d_sw_adjoint_case(fns_a, M, N, t, {n: f[n][t, 0] for n in f}, {n: sd[n][t, 0] for n in sd}, prm, 450.0, hydro, d_con)
            for k, v in r.items():
                out["%s.t%d.%s" % (name, t, k)] = v
        print("d_sw_fwd / d_sw_bwd", name, "done")
    import json
    out["cases"] = json.dumps({k: dict(prm=v[0], hydrostatic=v[1], d_con=v[2]) for k, v in cases.items()})
    np.savez_compressed(os.path.join(HERE, "ref_d_sw_tlm.npz"), **out)


def dyn_core_inputs(beta=0.0):
    """the state and two-sided configuration of tests/test_nh.py::test_dyn_core_nh_two_sided_emu and a seeded perturbation;
    beta > 0: three acoustic sub-steps with split_p_grad (the blended gradient du / dv is handed on twice)"""
    from common import rnd
    from test_nh import nh_state, NHCFG
    from test_dyn_core import CFG, TWO_SIDED
    from test_fv_dynamics import eta
    N, K = 12, 4
    ak, bk = eta(K, CFG["ptop"])
    f, rng = nh_state(N, K, 17, ak, bk)
    cfg = dict(NHCFG); cfg.update(n_split=2, bdt=600.0, a_imp=1.0); cfg.update(TWO_SIDED)
    if beta > 0.0:
        cfg.update(n_split=3, beta=beta)
    act = ["u", "v", "pt", "delp", "w", "delz"]
    d = {n: (1e-3 * np.abs(f[n]).mean() * rnd(rng, N, f[n].shape[1]) if n in act else 0 * f[n]) for n in f}
    return N, K, ak, bk, f, d, cfg, act


def main_dyn_core(beta=0.0):
    """DYN_CORE_TLM (non-hydrostatic, two acoustic sub-steps, six tiles with halo exchanges): tests/golden/ref_dyn_core_nh_tlm.npz;
    beta = 0.4 (SPLIT_P_GRAD_TLM instead of NH_P_GRAD_TLM, three sub-steps): tests/golden/ref_dyn_core_nh_beta_tlm.npz"""
    import ref_dyn_core as rd
    from common import metrics
    N, K, ak, bk, f, d, cfg, act = dyn_core_inputs(beta)
    ex = rd.Exchanger(N)
    consts = dict(rdgas=cfg["rdgas"], cp_air=cfg["cp_air"], grav=cfg["grav"])
    sp, fns, src = rd.load_reference(ex, consts, great_circle_dist, N)
    out = rd.run(fns, ex, grid_structs, metrics(N), N, K, f, d, cfg, ak, bk)
    np.savez_compressed(os.path.join(HERE, "ref_dyn_core_nh_beta_tlm.npz" if beta > 0.0 else "ref_dyn_core_nh_tlm.npz"),
                        **{k: v.astype(np.float64) for k, v in out.items()})
    print("dyn_core_tlm done, beta =", beta)


def dyn_core_hydro_inputs(beta=0.0, d_ext=0.0):
    """the state of tests/test_dyn_core.py (K = 5) with the two-sided configuration, three sub-steps, and a seeded perturbation"""
    from common import rnd
    from test_dyn_core import CFG, TWO_SIDED, hydro_state
    N, K = 12, 5
    f, rng = hydro_state(N, K, 5)
    cfg = dict(CFG); cfg.update(n_split=3, bdt=900.0, hydrostatic=True, rdgas=8314.47 / 28.965, grav=9.80665, p_fac=0.05); cfg.update(TWO_SIDED)
    if beta > 0.0:
        cfg["beta"] = beta
    if d_ext > 0.0:
        cfg["d_ext"] = d_ext
    act = ["u", "v", "pt", "delp"]
    d = {n: (1e-3 * np.abs(f[n]).mean() * rnd(rng, N, f[n].shape[1]) if n in act else 0 * f[n]) for n in f}
    bk = np.linspace(0.0, 1.0, K + 1); ak = CFG["ptop"] * (1.0 - bk)
    return N, K, ak, bk, f, d, cfg, act


def hydro_file(beta, d_ext):
    return "ref_dyn_core_hydro%s%s_tlm.npz" % ("_beta" if beta > 0.0 else "", "_dext" if d_ext > 0.0 else "")


def main_dyn_core_hydro(beta=0.0, d_ext=0.0):
    """DYN_CORE_TLM, hydrostatic (GEOPK_TLM, ONE_GRAD_P_TLM; beta = 0.4: GRAD1_P_UPDATE_TLM), three sub-steps, six tiles:
    tests/golden/ref_dyn_core_hydro[_beta]_tlm.npz"""
    import ref_dyn_core as rd
    from common import metrics
    N, K, ak, bk, f, d, cfg, act = dyn_core_hydro_inputs(beta, d_ext)
    ex = rd.Exchanger(N)
    consts = dict(rdgas=cfg["rdgas"], cp_air=cfg["cp_air"], grav=cfg["grav"])
    sp, fns, src = rd.load_reference(ex, consts, great_circle_dist, N)
    out = rd.run(fns, ex, grid_structs, metrics(N), N, K, f, d, cfg, ak, bk)
    np.savez_compressed(os.path.join(HERE, hydro_file(beta, d_ext)), **{k: v.astype(np.float64) for k, v in out.items()})
    print("dyn_core_tlm (hydrostatic) done, beta =", beta, "d_ext =", d_ext)


ADJ_OUT = ["u", "v", "pt", "delp", "w", "delz", "mfx", "cx"]          # the outputs of the library's dyn_core_nh module (tests/test_nh.py)
# outputs that carry a seed.  With seeds on u / w / delz (the outputs behind the LAST Riem_Solver3 of the loop) the reference's reverse sweep
# is not the transpose of its own tangent-linear code, see tests/test_ref_golden.py::test_oracle_reproduces_reference_dyn_core_adjoint
ADJ_SEEDED = ["pt", "delp", "mfx", "cx"]


def dyn_core_adjoint_inputs():
    """one-sided variant of dyn_core_inputs() (both sides run the trajectory's switches) and seeded output adjoints, scaled by the size of
    each output so that every path contributes to the input adjoints"""
    N, K, ak, bk, f, d, cfg, act = dyn_core_inputs()
    one = dict(cfg["traj"])
    cfg.update(one); cfg["traj"] = dict(one); cfg["split_damp"] = False; cfg["d2_bg_ks"] = cfg["d2_bg_k2"]
    npx = N + 1; C = (1, N, 1, N)
    regions = dict(u=(1, N, 1, npx), v=(1, npx, 1, N), pt=C, delp=C, w=C, delz=C, mfx=(1, npx, 1, N), cx=(1, npx, -2, N + 3))
    scale = np.load(os.path.join(HERE, "ref_dyn_core_nh_tlm.npz"))
    rng = np.random.default_rng(777)
    seed = {}
    for n in ADJ_OUT:
        i0, i1, j0, j1 = regions[n]
        s = np.zeros((6, K, N + 7, N + 7))
        if n in ADJ_SEEDED:
            s[:, :, j0 + 2: j1 + 3, i0 + 2: i1 + 3] = rng.standard_normal((6, K, j1 - j0 + 1, i1 - i0 + 1)) / np.abs(scale[n]).max()
        seed[n] = s
    return N, K, ak, bk, f, cfg, act, seed, regions


def main_dyn_core_adjoint():
    """DYN_CORE_FWD + DYN_CORE_BWD: the reference's reverse sweep of the non-hydrostatic acoustic loop (two sub-steps, six tiles, the adjoint
    halo exchanges served by the transposes of the cubed-sphere index maps): tests/golden/ref_dyn_core_nh_adm.npz"""
    import ref_dyn_core as rd
    from common import metrics
    N, K, ak, bk, f, cfg, act, seed, regions = dyn_core_adjoint_inputs()
    ex = rd.Exchanger(N)
    consts = dict(rdgas=cfg["rdgas"], cp_air=cfg["cp_air"], grav=cfg["grav"])
    sp, fns, src = rd.load_reference_adjoint(ex, consts, great_circle_dist, N)
    sd = dict(seed); sd["regions"] = regions
    out = rd.run_adjoint(fns, ex, grid_structs, metrics(N), N, K, f, sd, cfg, ak, bk)
    np.savez_compressed(os.path.join(HERE, "ref_dyn_core_nh_adm.npz"), **out)
    print("dyn_core_fwd / dyn_core_bwd done", {k: float(np.abs(v).max()) for k, v in out.items()})


FV_ACT = ["u", "v", "t", "delp", "qv", "ql", "qi", "o3", "w", "delz"]


def fv_act(hydro):
    return FV_ACT[:8] if hydro else FV_ACT


def fv_dynamics_inputs(K=5, hydro=False):
    """state and two-sided configuration of tests/test_step_api.py::test_step_api_two_sided_emu at npz = 5 (the reference skips the vertical
    remap for npz <= 4, fv_dynamics_tlm.F90: `IF (npz .GT. 4)`), and a seeded perturbation of the compute-domain prognostics"""
    from test_fv_dynamics import api_state, eta, ZVIR, RD
    from test_dyn_core import CFG, TWO_SIDED
    N = 12
    ak, bk = eta(K, CFG["ptop"])
    f, _ = api_state(N, K, 41, ak, bk, not hydro)
    cfg = dict(CFG); cfg.update(zvir=ZVIR, hydrostatic=hydro, k_split=1, n_split=2, dt=900.0, hord_tr=2, rdgas=RD, grav=9.80665, p_fac=0.05)
    cfg.update(TWO_SIDED)
    if hydro:                      # the reference's default external-mode damping (model/fv_arrays_nlm.F90:328), active in hydrostatic mode only
        cfg["d_ext"] = 0.02
    rng = np.random.default_rng(20261019)
    d = {}
    for n in fv_act(hydro):
        x = np.zeros_like(f[n])
        x[..., 3:3 + N, 3:3 + N] = 1e-3 * np.abs(f[n]).mean() * rng.standard_normal((6, K, N, N))
        d[n] = x
    return N, K, ak, bk, f, d, cfg


def main_fv_dynamics(hydro=False):
    """FV_DYNAMICS_TLM: one whole non-hydrostatic dynamics step (acoustic loop, tracer transport, vertical remap) on six tiles:
    tests/golden/ref_fv_dynamics_tlm.npz (compute-domain values and tangents of the ten prognostics)"""
    import torch
    import ref_dyn_core as rd
    import ref_fv_dynamics as rf
    from common import metrics, ograd
    from oracle.dyn_core import halo_of
    N, K, ak, bk, f, d, cfg = fv_dynamics_inputs(hydro=hydro)
    ACT = fv_act(hydro)
    g = ograd(N)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    st, st_tl = torch.func.jvp(lambda *a: rf.pre(dict(zip(ACT, a)), g, cfg, None), tuple(T(f[n]) for n in ACT), tuple(T(d[n]) for n in ACT))
    st = {k: v.numpy() for k, v in st.items()}; st_tl = {k: v.numpy() for k, v in st_tl.items()}
    phis = halo_of(N)[0].scalar(T(f["phis"])).numpy()
    ex = rd.Exchanger(N)
    consts = dict(rdgas=cfg["rdgas"], cp_air=cfg["cp_air"], grav=cfg["grav"], rvgas=461.5, pi=np.pi, radius=6371.0e3, kappa=cfg["akap"],
                  cp_vapor=4 * 461.5, hlv=2.5e6)
    sp, fns, src = rf.load_reference(ex, consts, great_circle_dist, N)
    out = rf.run(fns, ex, grid_structs, metrics(N), N, K, st, st_tl, phis, cfg, ak, bk)
    key = dict(t="pt")
    sav = {}
    for n in ACT:
        for sfx in ("", "_tl"):
            sav[n + sfx] = np.ascontiguousarray(out[key.get(n, n) + sfx][:, :, 3:3 + N, 3:3 + N])
    np.savez_compressed(os.path.join(HERE, "ref_fv_dynamics_hydro_tlm.npz" if hydro else "ref_fv_dynamics_tlm.npz"), **sav)
    print("fv_dynamics_tlm done, hydrostatic =", hydro)


def fv_dynamics_adjoint_inputs():
    """the hydrostatic whole-step case of fv_dynamics_inputs() with one-sided switches (both sides run the trajectory's; d_ext = 0.02) and
    seeded compute-domain output adjoints scaled by the size of each output"""
    N, K, ak, bk, f, d, cfg = fv_dynamics_inputs(hydro=True)
    one = dict(cfg["traj"])
    cfg.update(one); cfg["traj"] = dict(one); cfg["split_damp"] = False; cfg["d2_bg_ks"] = cfg["d2_bg_k2"]
    act = fv_act(True)
    scale = np.load(os.path.join(HERE, "ref_fv_dynamics_hydro_tlm.npz"))
    rng = np.random.default_rng(99)
    seed = {n: rng.standard_normal((6, K, N, N)) / np.abs(scale[n]).max() for n in act}
    return N, K, ak, bk, f, cfg, act, seed


def main_fv_dynamics_adjoint():
    """FV_DYNAMICS_FWD + FV_DYNAMICS_BWD: the reference's reverse sweep of one whole hydrostatic dynamics step (acoustic loop with the
    external-mode damping, tracer transport, vertical remap) on six tiles: tests/golden/ref_fv_dynamics_hydro_adm.npz (adjoints of the
    eight API prognostics on the compute domain, pulled back through the state conversion in front of fv_dynamics)"""
    import torch
    import ref_dyn_core as rd
    import ref_fv_dynamics as rf
    from common import metrics, ograd
    from oracle.dyn_core import halo_of
    N, K, ak, bk, f, cfg, act, seed = fv_dynamics_adjoint_inputs()
    g = ograd(N)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    x0 = tuple(T(f[n]) for n in act)
    st = rf.pre(dict(zip(act, x0)), g, cfg, None)
    keys = list(st)
    phis = halo_of(N)[0].scalar(T(f["phis"])).numpy()
    ex = rd.Exchanger(N)
    consts = dict(rdgas=cfg["rdgas"], cp_air=cfg["cp_air"], grav=cfg["grav"], rvgas=461.5, pi=np.pi, radius=6371.0e3, kappa=cfg["akap"],
                  cp_vapor=4 * 461.5, hlv=2.5e6)
    sp, fns, src = rf.load_reference_adjoint(ex, consts, great_circle_dist, N)
    st_ad = rf.run_adjoint(fns, ex, grid_structs, metrics(N), N, K, {k: v.numpy() for k, v in st.items()}, seed, act, phis, cfg, ak, bk)
    _, vjp_pre = torch.func.vjp(lambda *a: tuple(rf.pre(dict(zip(act, a)), g, cfg, None)[k] for k in keys), *x0)
    ad = vjp_pre(tuple(T(st_ad[k]) for k in keys))
    sav = {n + "_ad": np.ascontiguousarray(a.numpy()[:, :, 3:3 + N, 3:3 + N]) for n, a in zip(act, ad)}
    np.savez_compressed(os.path.join(HERE, "ref_fv_dynamics_hydro_adm.npz"), **sav)
    print("fv_dynamics_fwd / fv_dynamics_bwd done", {k: float(np.abs(v).max()) for k, v in sav.items()})


if __name__ == "__main__":
    if len(sys.argv) < 2 or sys.argv[1] == "d_sw":
        main()
    if len(sys.argv) < 2 or sys.argv[1] == "dyn_core":
        main_dyn_core()
    if len(sys.argv) < 2 or sys.argv[1] == "dyn_core_beta":
        main_dyn_core(0.4)
    if len(sys.argv) < 2 or sys.argv[1] == "dyn_core_hydro":
        main_dyn_core_hydro()
    if len(sys.argv) < 2 or sys.argv[1] == "dyn_core_hydro_beta":
        main_dyn_core_hydro(0.4)
    if len(sys.argv) < 2 or sys.argv[1] == "dyn_core_hydro_dext":          # external-mode damping (A2B_ORD2_TLM, the wk1 / wk2 terms of ONE_GRAD_P_TLM)
        main_dyn_core_hydro(0.0, 0.02)
    if len(sys.argv) < 2 or sys.argv[1] == "dyn_core_hydro_beta_dext":     # ... and of GRAD1_P_UPDATE_TLM
        main_dyn_core_hydro(0.4, 0.02)
    if len(sys.argv) < 2 or sys.argv[1] == "dyn_core_adjoint":
        main_dyn_core_adjoint()
    if len(sys.argv) < 2 or sys.argv[1] == "fv_dynamics":
        main_fv_dynamics()
    if len(sys.argv) < 2 or sys.argv[1] == "fv_dynamics_hydro":
        main_fv_dynamics(True)
    if len(sys.argv) < 2 or sys.argv[1] == "fv_dynamics_adjoint":
        main_fv_dynamics_adjoint()
