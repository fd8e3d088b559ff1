"""Probes behind the statements of DESIGN section 7 about the reference's reverse-mode routines (needs /root/reference: run in the build
container only, like make_ref_golden.py; nothing here is imported by the tests):
    python tests/golden/ref_adjoint_probes.py sim1     SIM1_SOLVER_FWD/BWD against SIM1_SOLVER_TLM      (dot-product identity per input)
    python tests/golden/ref_adjoint_probes.py riem3    RIEM_SOLVER3_FWD/BWD against RIEM_SOLVER3_TLM
    python tests/golden/ref_adjoint_probes.py c_sw     C_SW_FWD/BWD against torch.func.vjp of the oracle's c_sw, with an incoming w adjoint
All three are exact transposes (1e-15); the deviation of the reference's w adjoint comes from dyn_core_adm.F90 never assigning
nord_w_pert / damp_w_pert (tests/test_ref_golden.py::test_oracle_reproduces_reference_dyn_core_adjoint)."""
import os
import sys
import types
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p_ in (HERE, os.path.join(ROOT, "tests"), ROOT, os.path.join(ROOT, "fv3-jedi-linearmodel_b200")):
    sys.path.insert(0, p_)
import numpy as np
from ref_tlm import f90py
from ref_tlm.f90py import FA
REF = "/root/reference/src/dynamics/atmos_cubed_sphere/model_tlmadm/"
consts = dict(rdgas=287.05, cp_air=1004.6, grav=9.80665)


def probe_sim1():
    _, fa_, _ = f90py.load([REF + "nh_utils_adm.F90"], extra=dict(consts), strict=False, only={"sim1_solver_fwd", "sim1_solver_bwd", "sim1_solver"})
    _, ft_, _ = f90py.load([REF + "nh_utils_tlm.F90"], extra=dict(consts), strict=False, only={"sim1_solver_tlm", "sim1_solver"})
    print(list(fa_), list(ft_))
    rng = np.random.default_rng(3)
    km, n = 5, 2
    def A(x): return FA(np.asfortranarray(x.copy()), (1, 1)) if x.ndim == 2 else FA(x.copy(), (1,))
    dm2 = 1000.0 + 100 * rng.random((n, km)); pt2 = 300 + 10 * rng.random((n, km)); dz2 = -(800 + 100 * rng.random((n, km))); w2 = rng.standard_normal((n, km))
    pm2 = 5e4 + 1e3 * rng.random((n, km)); pem = np.cumsum(np.concatenate([np.full((n, 1), 100.0), dm2 * 9.8], axis=1), axis=1)
    gm2 = np.zeros((n, km)); cp2 = np.zeros((n, km)); ws = 0.01 * rng.standard_normal(n)
    dt, rgas, gama, kappa, p_fac = 300.0, 287.05, 1.4, 2. / 7., 0.05
    names = ["dm2", "pm2", "pem", "w2", "dz2", "pt2", "ws"]
    base = dict(dm2=dm2, pm2=pm2, pem=pem, w2=w2, dz2=dz2, pt2=pt2, ws=ws)
    def tlm(d):
        v = {k: A(x) for k, x in base.items()}; t = {k: A(d.get(k, np.zeros_like(base[k]))) for k in base}
        pe = FA.alloc(((1, n), (1, km + 1))); pe_tl = FA.alloc(((1, n), (1, km + 1)))
        ft_["sim1_solver_tlm"](dt, 1, n, km, rgas, gama, A(gm2), A(cp2), kappa, pe, pe_tl, v["dm2"], t["dm2"], v["pm2"], t["pm2"], v["pem"], t["pem"], v["w2"], t["w2"], v["dz2"], t["dz2"], v["pt2"], t["pt2"], v["ws"], t["ws"], p_fac)
        return dict(pe=pe_tl.a.copy(), w2=t["w2"].a.copy(), dz2=t["dz2"].a.copy())
    def adj(seed):
        v = {k: A(x) for k, x in base.items()}
        pe = FA.alloc(((1, n), (1, km + 1)))
        assert not f90py.stack()
        fa_["sim1_solver_fwd"](dt, 1, n, km, rgas, gama, A(gm2), A(cp2), kappa, pe, v["dm2"], v["pm2"], v["pem"], v["w2"], v["dz2"], v["pt2"], v["ws"], p_fac)
        a = {k: A(np.zeros_like(base[k])) for k in base}
        a["w2"] = A(seed["w2"]); a["dz2"] = A(seed["dz2"]); pe_ad = A(seed["pe"])
        fa_["sim1_solver_bwd"](dt, 1, n, km, rgas, gama, A(gm2), A(cp2), kappa, pe, pe_ad, v["dm2"], a["dm2"], v["pm2"], a["pm2"], v["pem"], a["pem"], v["w2"], a["w2"], v["dz2"], a["dz2"], v["pt2"], a["pt2"], v["ws"], a["ws"], p_fac)
        assert not f90py.stack()
        return {k: a[k].a.copy() for k in base}
    seed = dict(pe=rng.standard_normal((n, km + 1)), w2=rng.standard_normal((n, km)), dz2=rng.standard_normal((n, km)))
    ad = adj(seed)
    for k in names:
        d = {k: rng.standard_normal(base[k].shape)}
        o = tlm(d)
        lhs = sum(float((o[q] * seed[q]).sum()) for q in seed)
        rhs = float((d[k] * ad[k]).sum())
        print(k, "dot <M dx, y> = %.12e  <dx, MT y> = %.12e  rel %.1e" % (lhs, rhs, abs(lhs - rhs) / max(abs(lhs), 1e-300)))


def probe_riem3():
    _, fa_, _ = f90py.load([REF + "nh_core_adm.F90", REF + "nh_utils_adm.F90"], extra=dict(consts), strict=False)
    _, ft_, _ = f90py.load([REF + "nh_core_tlm.F90", REF + "nh_utils_tlm.F90"], extra=dict(consts), strict=False)
    rng = np.random.default_rng(3)
    km, N, ng = 5, 3, 3
    isd, ied = 1 - ng, N + ng
    n = ied - isd + 1
    B3 = lambda nk: ((isd, ied), (isd, ied), (1, nk))
    bnd = dict(w=B3(km), delz=B3(km), pt=B3(km), delp=B3(km), zh=B3(km + 1), pe=((0, N + 1), (1, km + 1), (0, N + 1)), ppe=B3(km + 1), pk3=B3(km + 1),
               pk=((1, N), (1, N), (1, km + 1)), peln=((1, N), (1, km + 1), (1, N)), ws=((1, N), (1, N)))
    shape = lambda b: tuple(hi - lo + 1 for lo, hi in b)
    base = {}
    base["delp"] = 2000.0 + 200 * rng.random(shape(bnd["delp"])); base["pt"] = 300 + 10 * rng.random(shape(bnd["pt"]))
    dz = -(800 + 100 * rng.random(shape(bnd["delz"]))); base["delz"] = dz
    zh = np.zeros(shape(bnd["zh"])); zh[:, :, km] = 10.0
    for k in range(km - 1, -1, -1): zh[:, :, k] = zh[:, :, k + 1] - dz[:, :, k]
    base["zh"] = zh; base["w"] = rng.standard_normal(shape(bnd["w"])); base["ws"] = 0.01 * rng.standard_normal(shape(bnd["ws"]))
    for k in ("pe", "ppe", "pk3", "pk", "peln"): base[k] = np.zeros(shape(bnd[k]))
    zs = FA(np.full((n, n), 10.0), (isd, isd)); q_con = FA.alloc(B3(km)); cappa = FA.alloc(B3(km))
    def F(k, x): return FA(np.asfortranarray(x.copy()), tuple(lo for lo, hi in bnd[k]))
    order = ["w", "delz", "pt", "delp", "zh", "pe", "ppe", "pk3", "pk", "peln", "ws"]
    args0 = (0, 300.0, 1, N, 1, N, km, ng, isd, ied, isd, ied, 2. / 7., cappa, 1004.6, 100.0, zs, q_con)
    tail = (0.0, 0.05, 1.0, False, False, False)
    def tlm(d):
        v = {k: F(k, base[k]) for k in order}; t = {k: F(k, d.get(k, np.zeros_like(base[k]))) for k in order}
        pairs = []
        for k in order: pairs += [v[k], t[k]]
        ft_["riem_solver3_tlm"](*args0, *pairs, *tail)
        return {k: t[k].a.copy() for k in order}
    def adj(seed):
        v = {k: F(k, base[k]) for k in order}
        assert not f90py.stack()
        fa_["riem_solver3_fwd"](*args0, *[v[k] for k in order], *tail)
        a = {k: F(k, seed.get(k, np.zeros_like(base[k]))) for k in order}
        pairs = []
        for k in order: pairs += [v[k], a[k]]
        fa_["riem_solver3_bwd"](*args0, *pairs, *tail)
        assert not f90py.stack()
        return {k: a[k].a.copy() for k in order}
    C = (slice(ng, ng + N), slice(ng, ng + N))
    seed = {}
    for k in ("w", "delz", "zh", "ppe"):
        s = np.zeros_like(base[k]); s[C] = rng.standard_normal(s[C].shape); seed[k] = s
    ad = adj(seed)
    for k in ("w", "delz", "pt", "delp", "zh", "ws"):
        dd = np.zeros_like(base[k])
        if k == "ws": dd = rng.standard_normal(dd.shape)
        else: dd[C] = rng.standard_normal(dd[C].shape)
        o = tlm({k: dd})
        lhs = sum(float((o[q] * seed[q]).sum()) for q in seed)
        rhs = float((dd * ad[k]).sum())
        print(k, "dot <M dx, y> = %.12e  <dx, MT y> = %.12e  rel %.1e" % (lhs, rhs, abs(lhs - rhs) / max(abs(lhs), 1e-300)))


def probe_c_sw():
    import torch
    import make_ref_golden as m
    from common import metrics, ograd, rnd, relerr, region
    from oracle import sw_core as osw
    from test_c_sw import smooth_state
    extra = dict(ng=3, great_circle_dist=m.great_circle_dist, fpp=types.SimpleNamespace(fpp_overload_r4=False))
    _, fns, _ = f90py.load([REF + "tp_core_adm.F90", REF + "sw_core_adm.F90", REF + "a2b_edge_adm.F90"], extra=extra, strict=False)
    N, K = 12, 1
    f, rng = smooth_state(N, K, 71)
    M = metrics(N); g = ograd(N)
    dt2 = 225.0; t = 1
    names = ["delp", "pt", "u", "v", "w"]
    onames = ["delpc", "ptc", "uc", "vc", "ua", "va", "ut", "vt", "divg_d", "wc"]
    npx = N + 1
    regs = {"delpc": (0, npx, 0, npx), "ptc": (0, npx, 0, npx), "wc": (0, npx, 0, npx), "uc": (1, npx, 1, N), "vc": (1, N, 1, npx),
            "ua": (0, npx, 0, npx), "va": (0, npx, 0, npx), "ut": (0, npx + 1, 0, npx), "vt": (0, npx, 0, npx + 1), "divg_d": (1, npx, 1, npx)}
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    seeds = {}
    for n in onames:
        s = np.zeros((6, K, N + 7, N + 7)); i0, i1, j0, j1 = regs[n]
        s[t, 0, j0 + 2: j1 + 3, i0 + 2: i1 + 3] = rng.standard_normal((j1 - j0 + 1, i1 - i0 + 1))
        seeds[n] = s
    w_in_ad = np.zeros((6, K, N + 7, N + 7)); w_in_ad[t, 0, 3:N + 3, 3:N + 3] = rng.standard_normal((N, N))
    def fn(*a):
        o = osw.c_sw(*a, g, dt2, False, 1)
        return tuple(o[k] for k in onames) + (a[4],)        # w passes through
    _, vjp = torch.func.vjp(fn, *[T(f[n]) for n in names])
    ad_o = vjp(tuple(T(seeds[n]) for n in onames) + (T(w_in_ad),))
    # reference
    bd, gs, fl = m.grid_structs(M, t, N)
    isd, ied, jsd, jed = bd.isd, bd.ied, bd.jsd, bd.jed
    A = (isd, ied, jsd, jed)
    bnd = dict(delp=A, pt=A, w=A, ua=A, va=A, delpc=A, ptc=A, wc=A, ut=A, vt=A, u=(isd, ied, jsd, jed + 1), vc=(isd, ied, jsd, jed + 1), v=(isd, ied + 1, jsd, jed),
               uc=(isd, ied + 1, jsd, jed), divg_d=(isd, ied + 1, jsd, jed + 1))
    Z = lambda b: FA.alloc(((b[0], b[1]), (b[2], b[3])))
    a = {n: (m.fa(f[n][t, 0], *bnd[n]) if n in names else Z(bnd[n])) for n in bnd}
    order = ["delpc", "delp", "ptc", "pt", "u", "v", "w", "uc", "vc", "ua", "va", "wc", "ut", "vt", "divg_d"]
    assert not f90py.stack()
    fns["c_sw_fwd"](*[a[n] for n in order], 1, dt2, False, True, bd, gs, fl)
    ad = {n: (m.fa(seeds[n][t, 0], *bnd[n]) if n in onames else Z(bnd[n])) for n in bnd}
    ad["w"] = m.fa(w_in_ad[t, 0], *bnd["w"])
    pairs = []
    for n in order: pairs += [a[n], ad[n]]
    fns["c_sw_bwd"](*pairs, 1, dt2, False, True, bd, gs, fl)
    assert not f90py.stack()
    for n, o in zip(names, ad_o):
        r = m.back(ad[n], N)
        print(n, "relerr", relerr(o[t, 0].numpy(), r), "max", np.abs(r).max())
    r = m.back(ad["w"], N); o = ad_o[4][t, 0].numpy(); wi = w_in_ad[t, 0]
    d = r - o
    np.set_printoptions(linewidth=250, precision=2, suppress=True)
    print("diff / incoming (where incoming != 0):")
    mask = wi != 0
    ratio = np.zeros_like(d); ratio[mask] = d[mask] / wi[mask]
    print(ratio[0:N + 7, 0:N + 7])


if __name__ == "__main__":
    {"sim1": probe_sim1, "riem3": probe_riem3, "c_sw": probe_c_sw}[sys.argv[1]]()
