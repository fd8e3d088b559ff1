"""ORACLE (test infrastructure only).  Vertical Lagrangian-to-Eulerian remap restated in torch
float64 from model/fv_mapz_nlm.F90  Lagrangian_to_Eulerian :60-958 (remap_option 0: T in
log p), map_scalar :1237, map1_ppm :1332, map1_q2 :1541, cs_profile :2113 / scalar_profile
:1730 for the un-limited |kord| > 16 branch (the only one the TL/AD implement,
model_tlmadm/fv_mapz_tlm.F90:8493-8507, :8652-8666).

The layer search of the reference (monotone cursor with GOTOs, :1288-1327) is restated as
explicit overlap masks between target layer k and source layer m; every per-piece formula
is the reference's.  Summation over the source layers uses torch.sum (pairwise) instead of
the sequential loop: differences are O(1e-16) relative.

parity: pinned through the reference's own LAGRANGIAN_TO_EULERIAN_TLM / _FWD / _BWD executed inside FV_DYNAMICS_TLM (non-hydrostatic and
hydrostatic) and FV_DYNAMICS_FWD/BWD (hydrostatic): tests/test_ref_golden.py, fixtures tests/golden/ref_fv_dynamics_*.npz.  Arrays [6, K, NY, NX] / [6, K+1, NY, NX].
"""
import numpy as np
import torch

R3, R23 = 1. / 3., 2. / 3.


def cs_profile(a, dp, iv=0, qs=None):
    """interface values q(1..K+1) of the cubic-spline profile, |kord| > 16.
    a, dp: [6, K, ...].  iv = -2: w with bottom value qs.  Returns [6, K+1, ...]."""
    K = a.shape[1]
    A = lambda k: a[:, k - 1]
    D = lambda k: dp[:, k - 1]
    q = [None] * (K + 2)
    gam = [None] * (K + 2)
    if iv == -2:
        gam[2] = 0.5 + 0.0 * A(1)
        q[1] = 1.5 * A(1)
        for k in range(2, K):
            grat = D(k - 1) / D(k)
            bet = 2. + grat + grat - gam[k]
            q[k] = (3. * (A(k - 1) + A(k)) - q[k - 1]) / bet
            gam[k + 1] = grat / bet
        grat = D(K - 1) / D(K)
        q[K] = (3. * (A(K - 1) + A(K)) - grat * qs - q[K - 1]) / (2. + grat + grat - gam[K])
        q[K + 1] = qs + 0.0 * A(1)
        for k in range(K - 1, 0, -1):
            q[k] = q[k] - gam[k + 1] * q[k + 1]
    else:
        grat = D(2) / D(1)
        bet = grat * (grat + 0.5)
        q[1] = ((grat + grat) * (grat + 1.) * A(1) + A(2)) / bet
        gam[1] = (1. + grat * (grat + 1.5)) / bet
        d4 = None
        for k in range(2, K + 1):
            d4 = D(k - 1) / D(k)
            bet = 2. + d4 + d4 - gam[k - 1]
            q[k] = (3. * (A(k - 1) + d4 * A(k)) - q[k - 1]) / bet
            gam[k] = d4 / bet
        a_bot = 1. + d4 * (d4 + 1.5)
        q[K + 1] = (2. * d4 * (d4 + 1.) * A(K) + A(K - 1) - a_bot * q[K]) / (d4 * (d4 + 0.5) - a_bot * gam[K])
        for k in range(K, 0, -1):
            q[k] = q[k] - gam[k] * q[k + 1]
    return torch.stack(q[1:K + 2], dim=1)


def cs_limiters(extm, a, a2, a3, a4, iv):
    """model/fv_mapz_nlm.F90:2467-2542 on tensors (one layer)"""
    z = torch.zeros_like(a)
    if iv == 0:
        safe = torch.where(a4 == 0., torch.ones_like(a4), a4)
        neg = ((a3 - a2).abs() < -a4) & ((a + 0.25 * (a3 - a2) ** 2 / safe + a4 * (1. / 12.)) < 0.)
        c1 = (a < a3) & (a < a2)
        c2 = a3 > a2
        a4n = torch.where(c1, z, torch.where(c2, 3. * (a2 - a), 3. * (a3 - a)))
        a3n = torch.where(c1, a, torch.where(c2, a2 - a4n, a3))
        a2n = torch.where(c1, a, torch.where(c2, a2, a3 - a4n))
        a2o = torch.where(neg, a2n, a2); a3o = torch.where(neg, a3n, a3); a4o = torch.where(neg, a4n, a4)
        le = a <= 0.
        return torch.where(le, a, a2o), torch.where(le, a, a3o), torch.where(le, z, a4o)
    flat = ((a - a2) * (a - a3) >= 0.) if iv == 1 else extm
    da1 = a3 - a2; da2 = da1 ** 2; a6da = a4 * da1
    lo = a6da < -da2; hi = (~lo) & (a6da > da2)
    a4n = torch.where(lo, 3. * (a2 - a), torch.where(hi, 3. * (a3 - a), a4))
    a3n = torch.where(lo, a2 - a4n, a3)
    a2n = torch.where(hi, a3 - a4n, a2)
    return torch.where(flat, a, a2n), torch.where(flat, a, a3n), torch.where(flat, z, a4n)


def limited_profile(a, qi, iv, kord, qmin=0., cs=False):
    """the monotone sub-grid profiles of the nonlinear model, |kord| = 8 .. 14: scalar_profile (:1814-2109; cs = False) and
    cs_profile (:2197-2463; cs = True, no qmin tests).  a [6,K,..] layer means, qi [6,K+1,..] spline interface values.
    Returns a42, a43, a44 [6,K,..]."""
    K = a.shape[1]
    ak = abs(kord)
    A = [a[:, k] for k in range(K)]
    q = [qi[:, k] for k in range(K + 1)]
    mx, mn = torch.maximum, torch.minimum
    q[1] = mn(q[1], mx(A[0], A[1])); q[1] = mx(q[1], mn(A[0], A[1]))
    gm = [None] + [A[k] - A[k - 1] for k in range(1, K)]
    for k in range(2, K - 1):
        lo, hi = mn(A[k - 1], A[k]), mx(A[k - 1], A[k])
        both = mx(mn(q[k], hi), lo)
        lmax = mx(q[k], lo)
        lmin = mn(q[k], hi)
        if iv == 0:
            lmin = mx(torch.zeros_like(lmin), lmin)
        q[k] = torch.where(gm[k - 1] * gm[k + 1] > 0., both, torch.where(gm[k - 1] > 0., lmax, lmin))
    q[K - 1] = mn(q[K - 1], mx(A[K - 2], A[K - 1])); q[K - 1] = mx(q[K - 1], mn(A[K - 2], A[K - 1]))
    A2 = [q[k] for k in range(K)]; A3 = [q[k + 1] for k in range(K)]; A4 = [None] * K
    extm = [((A2[k] - A[k]) * (A3[k] - A[k]) > 0.) if k in (0, K - 1) else (gm[k] * gm[k + 1] < 0.) for k in range(K)]
    f3 = lambda k: 3. * (2. * A[k] - (A2[k] + A3[k]))
    f6 = lambda k: 6. * A[k] - 3. * (A2[k] + A3[k])
    z = torch.zeros_like(A[0])
    def huynh(k):
        pmp_1 = A[k] - 2. * gm[k + 1]; lac_1 = pmp_1 + 1.5 * gm[k + 2]
        l = mn(mx(A2[k], mn(mn(A[k], pmp_1), lac_1)), mx(mx(A[k], pmp_1), lac_1))
        pmp_2 = A[k] + 2. * gm[k]; lac_2 = pmp_2 - 1.5 * gm[k - 1]
        r = mn(mx(A3[k], mn(mn(A[k], pmp_2), lac_2)), mx(mx(A[k], pmp_2), lac_2))
        return l, r
    if iv == 0:
        A2[0] = mx(z, A2[0])
    elif iv == -1:
        A2[0] = torch.where(A2[0] * A[0] <= 0., z, A2[0])
    A4[0] = f3(0); A2[0], A3[0], A4[0] = cs_limiters(extm[0], A[0], A2[0], A3[0], A4[0], 1)
    A4[1] = f3(1); A2[1], A3[1], A4[1] = cs_limiters(extm[1], A[1], A2[1], A3[1], A4[1], 2)
    for k in range(2, K - 2):
        small = (A[k] < qmin) if not cs else torch.zeros_like(extm[k])
        if ak < 9:
            A2[k], A3[k] = huynh(k); A4[k] = f3(k)
        elif ak in (9, 10, 12):
            f = (f6 if (cs or ak != 9) else f3)
            a4u = f(k)
            l, r = huynh(k)
            need = a4u.abs() > (A2[k] - A3[k]).abs()
            a2s = torch.where(need, l, A2[k]); a3s = torch.where(need, r, A3[k])
            a4s = (6. * A[k] - 3. * (a2s + a3s)) if f is f6 else 3. * (2. * A[k] - (a2s + a3s))
            if ak == 9:
                fl = extm[k] & (extm[k - 1] | extm[k + 1] | small)
                A2[k] = torch.where(fl, A[k], a2s); A3[k] = torch.where(fl, A[k], a3s); A4[k] = torch.where(fl, z, a4s)
            elif ak == 10:
                fl = extm[k] & (small | extm[k - 1] | extm[k + 1])
                ex = extm[k] & ~fl         # true local extremum: unlimited
                A2[k] = torch.where(fl, A[k], torch.where(ex, A2[k], a2s)); A3[k] = torch.where(fl, A[k], torch.where(ex, A3[k], a3s))
                A4[k] = torch.where(fl, z, torch.where(ex, a4u, a4s))
            else:
                fl = extm[k]
                A2[k] = torch.where(fl, A[k], a2s); A3[k] = torch.where(fl, A[k], a3s); A4[k] = torch.where(fl, z, a4s)
        elif ak == 13:
            fl = extm[k] & extm[k - 1] & extm[k + 1]
            l, r = huynh(k)
            lim = extm[k] & ~fl
            a2s = torch.where(lim, l, A2[k]); a3s = torch.where(lim, r, A3[k])
            A2[k] = torch.where(fl, A[k], a2s); A3[k] = torch.where(fl, A[k], a3s)
            A4[k] = torch.where(fl, z, 3. * (2. * A[k] - (a2s + a3s)))
        elif ak == 14:
            A4[k] = f3(k)
        else:       # 11
            fl = extm[k] & (extm[k - 1] | extm[k + 1] | small)
            a4u = f3(k)
            A2[k] = torch.where(fl, A[k], A2[k]); A3[k] = torch.where(fl, A[k], A3[k]); A4[k] = torch.where(fl, z, a4u)
        if iv == 0:
            A2[k], A3[k], A4[k] = cs_limiters(extm[k], A[k], A2[k], A3[k], A4[k], 0)
    if iv == 0:
        A3[K - 1] = mx(z, A3[K - 1])
    elif iv == -1:
        A3[K - 1] = torch.where(A3[K - 1] * A[K - 1] <= 0., z, A3[K - 1])
    A4[K - 2] = f3(K - 2); A2[K - 2], A3[K - 2], A4[K - 2] = cs_limiters(extm[K - 2], A[K - 2], A2[K - 2], A3[K - 2], A4[K - 2], 2)
    A4[K - 1] = f3(K - 1); A2[K - 1], A3[K - 1], A4[K - 1] = cs_limiters(extm[K - 1], A[K - 1], A2[K - 1], A3[K - 1], A4[K - 1], 1)
    return torch.stack(A2, 1), torch.stack(A3, 1), torch.stack(A4, 1)


def remap_integrate(a, qi, pe1, pe2, dp2=None, a4=None):
    """mean of the piecewise-parabolic profile (a4(1)=a, a4(2)=qi(k), a4(3)=qi(k+1),
    a4(4)=3(2a-(a4(2)+a4(3)))) of the source layers pe1 over each target layer pe2.
    map_scalar / map1_ppm divide by (pe2(k+1)-pe2(k)); map1_q2 by the given dp2."""
    K = a.shape[1]
    a41 = a
    if a4 is None:
        a42 = qi[:, :-1]; a43 = qi[:, 1:]
        a44 = 3. * (2. * a41 - (a42 + a43))
    else:
        a42, a43, a44 = a4
    dp1 = pe1[:, 1:] - pe1[:, :-1]
    # layer index (0-based) of each target interface: number of interior source edges strictly below it
    with torch.no_grad():
        cnt = (pe1[:, None, 1:K] < pe2[:, :, None]).sum(dim=2)          # [6, K+1(target edge), ...]
        l_top = cnt[:, :-1]                                                  # layer containing pe2(k)
        # layer containing pe2(k+1): the reference tests "pe2(k+1) <= pe1(l+1)" -> ties stay in the upper layer
        cnt2 = (pe1[:, None, 1:K] < pe2[:, 1:, None]).sum(dim=2)
        l_bot = cnt2
    idx = torch.arange(K).view(1, 1, K, *([1] * (a.dim() - 2)))          # source layer m along dim 2
    lt = l_top[:, :, None]; lb = l_bot[:, :, None]
    def src(x): return x[:, None]                                           # [6, 1, K(m), ...]
    p2k = pe2[:, :-1, None]; p2k1 = pe2[:, 1:, None]
    pl = (p2k - src(pe1[:, :-1])) / src(dp1)
    pr = (p2k1 - src(pe1[:, :-1])) / src(dp1)
    same = (lt == lb)
    # case 1: target layer inside one source layer
    one = src(a42) + 0.5 * (src(a44) + src(a43) - src(a42)) * (pr + pl) - src(a44) * R3 * (pr * (pr + pl) + pl ** 2)
    # case 2: first (partial) source layer
    first = (src(pe1[:, 1:]) - p2k) * (src(a42) + 0.5 * (src(a44) + src(a43) - src(a42)) * (1. + pl) - src(a44) * (R3 * (1. + pl * (1. + pl))))
    full = src(dp1) * src(a41)
    dpl = p2k1 - src(pe1[:, :-1])
    esl = dpl / src(dp1)
    last = dpl * (src(a42) + 0.5 * esl * (src(a43) - src(a42) + src(a44) * (1. - R23 * esl)))
    zero = torch.zeros_like(one)
    contrib = torch.where((idx == lt) & ~same, first, zero) + torch.where((idx > lt) & (idx < lb), full, zero) + \
        torch.where((idx == lb) & ~same, last, zero)
    qsum = contrib.sum(dim=2)
    den = (pe2[:, 1:] - pe2[:, :-1]) if dp2 is None else dp2
    one_v = torch.where((idx == lt) & same, one, zero).sum(dim=2)
    same2 = (l_top == l_bot)
    return torch.where(same2, one_v, qsum / den)


def remap(a, pe1, pe2, iv=0, qs=None, dp2=None, kord=17, qmin=0., cs=False, kord_pert=None):
    """kord: the mapping order (|kord| > 16: unlimited cubic spline; 8 .. 14: limited profiles of the nonlinear model).
    kord_pert (two-sided mode, split_kord, model_tlmadm/fv_mapz_tlm.F90:494-506): the increment is remapped with kord_pert
    about the same inputs, the trajectory with kord."""
    dp1 = pe1[:, 1:] - pe1[:, :-1]
    qi = cs_profile(a, dp1, iv, qs)
    def one(ko):
        if abs(ko) > 16:
            return remap_integrate(a, qi, pe1, pe2, dp2)
        return remap_integrate(a, qi, pe1, pe2, dp2, a4=limited_profile(a, qi, iv, ko, qmin, cs))
    if kord_pert is None or abs(kord_pert) == abs(kord):
        return one(kord)
    from .d_sw import splice
    return splice(one(kord_pert), one(kord))


def put(a, sl, v):
    a = a.clone(); a[sl] = v; return a


def lagrangian_to_eulerian(st, g, ak, bk, cfg, last_step):
    """hydrostatic / non-hydrostatic remap, remap_option = 0, consv = 0, no omega.
    st: dict with pe, pk, peln [6,K+1], delp, pt, u, v, q (list), [w, delz, ws]; pe must be
    valid on is-1..ie+1.  Returns the updated dict (pt = T_v/(1+zvir q) on the last step,
    theta_v otherwise, :883-931)."""
    from .cubed_sphere import R
    N = g.N
    is_, ie, js, je = 1, N, 1, N
    akap, ptop, zvir = cfg["akap"], cfg["ptop"], cfg["zvir"]
    hydro = cfg["hydrostatic"]
    K = st["delp"].shape[1]
    AK = torch.as_tensor(ak, dtype=torch.float64).view(1, -1, 1, 1)
    BK = torch.as_tensor(bk, dtype=torch.float64).view(1, -1, 1, 1)
    pe, pk, peln = st["pe"], st["pk"], st["peln"]
    C = (slice(None), slice(None), R(js, je), R(is_, ie))
    pe1 = pe[C]
    ps = pe1[:, K:K + 1]
    pe2 = torch.cat([0 * ps + ptop, AK[:, 1:K] + BK[:, 1:K] * ps, ps], dim=1)
    dp2 = pe2[:, 1:] - pe2[:, :-1]
    pk1 = pk[C]; pn1 = peln[C]
    pt = st["pt"][C]
    if hydro:
        pt = pt * (pk1[:, 1:] - pk1[:, :-1]) / (akap * (pn1[:, 1:] - pn1[:, :-1]))
    else:
        k1k = akap / (1. - akap); rrg = -cfg["rdgas"] / cfg["grav"]
        delz = st["delz"][C]; delp0 = st["delp"][C]
        pt = pt * torch.exp(k1k * torch.log(rrg * delp0 / delz * pt))
        delz = -delz / delp0
    pn2 = torch.cat([pn1[:, :1], torch.log(pe2[:, 1:K]), pn1[:, K:]], dim=1)
    pk2 = torch.cat([pk1[:, :1], torch.exp(akap * pn2[:, 1:K]), pk1[:, K:]], dim=1)
    tj = cfg.get("traj") or {}
    kp = 17 if tj else None                                   # the increment always uses the linear |kord| = 17
    k_mt, k_wz, k_tm, k_tr = (tj.get(n, 17) or 17 for n in ("kord_mt", "kord_wz", "kord_tm", "kord_tr"))
    pt = remap(pt, pn1, pn2, iv=1, kord=k_tm, qmin=184., kord_pert=kp)
    qn = [remap(q[C], pe1, pe2, iv=0, dp2=dp2, kord=k_tr, qmin=0., kord_pert=kp) for q in st["q"]]
    out = dict(st)
    if not hydro:
        w = remap(st["w"][C], pe1, pe2, iv=-2, qs=st["ws"][C][:, 0], kord=k_wz, cs=True, kord_pert=kp)
        delz = remap(delz, pe1, pe2, iv=1, kord=k_tm, cs=True, kord_pert=kp)
        delz = -delz * dp2
        out["w"] = put(st["w"], C, w); out["delz"] = put(st["delz"], C, delz)
        pkz = torch.exp(akap * torch.log(rrg * dp2 / delz * pt))
    else:
        pkz = (pk2[:, 1:] - pk2[:, :-1]) / (akap * (pn2[:, 1:] - pn2[:, :-1]))
    # u (rows js..je+1) and v (cols is..ie+1)
    Cu = (slice(None), slice(None), R(js, je + 1), R(is_, ie))
    Cum = (slice(None), slice(None), R(js - 1, je), R(is_, ie))
    pe0 = 0.5 * (pe[Cum] + pe[Cu])
    pe0 = torch.cat([pe[Cu][:, :1], pe0[:, 1:]], dim=1)
    pe3 = AK + 0.5 * BK * (pe[Cum][:, K:] + pe[Cu][:, K:])
    u = remap(st["u"][Cu], pe0, pe3, iv=-1, kord=k_mt, cs=True, kord_pert=kp)
    Cv = (slice(None), slice(None), R(js, je), R(is_, ie + 1))
    Cvm = (slice(None), slice(None), R(js, je), R(is_ - 1, ie))
    pe0 = 0.5 * (pe[Cvm] + pe[Cv])
    pe0 = torch.cat([pe[Cv][:, :1] * 0 + pe[Cv][:, :1], pe0[:, 1:]], dim=1)
    pe3 = AK + 0.5 * BK * (pe[Cvm][:, K:] + pe[Cv][:, K:])
    pe3 = torch.cat([0 * pe3[:, :1] + float(ak[0]), pe3[:, 1:]], dim=1)
    v = remap(st["v"][Cv], pe0, pe3, iv=-1, kord=k_mt, cs=True, kord_pert=kp)
    if last_step:
        pt = pt / (1. + zvir * qn[0])
    else:
        pt = pt / pkz
    out["pt"] = put(st["pt"], C, pt); out["delp"] = put(st["delp"], C, dp2)
    out["q"] = [put(q, C, n) for q, n in zip(st["q"], qn)]
    out["u"] = put(st["u"], Cu, u); out["v"] = put(st["v"], Cv, v)
    out["pe"] = put(pe, C, pe2); out["pk"] = put(pk, C, pk2); out["peln"] = put(peln, C, pn2)
    out["pkz"] = put(st["pkz"], C, pkz)
    return out
