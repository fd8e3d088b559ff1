"""ORACLE (test infrastructure only).  Vertical Lagrangian-to-Eulerian remap restated in torch
float64 from model/fv_mapz_nlm.F90  Lagrangian_to_Eulerian :60-958 (remap_option 0: T in
log p), map_scalar :1237, map1_ppm :1332, map1_q2 :1541, cs_profile :2113 / scalar_profile
:1730 for the un-limited |kord| > 16 branch (the only one the TL/AD implement,
model_tlmadm/fv_mapz_tlm.F90:8493-8507, :8652-8666).

The layer search of the reference (monotone cursor with GOTOs, :1288-1327) is restated as
explicit overlap masks between target layer k and source layer m; every per-piece formula
is the reference's.  Summation over the source layers uses torch.sum (pairwise) instead of
the sequential loop: differences are O(1e-16) relative.

parity unpinned (no reference vectors).  Arrays [6, K, NY, NX] / [6, K+1, NY, NX].
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


def remap_integrate(a, qi, pe1, pe2, dp2=None):
    """mean of the piecewise-parabolic profile (a4(1)=a, a4(2)=qi(k), a4(3)=qi(k+1),
    a4(4)=3(2a-(a4(2)+a4(3)))) of the source layers pe1 over each target layer pe2.
    map_scalar / map1_ppm divide by (pe2(k+1)-pe2(k)); map1_q2 by the given dp2."""
    K = a.shape[1]
    a41 = a; a42 = qi[:, :-1]; a43 = qi[:, 1:]
    a44 = 3. * (2. * a41 - (a42 + a43))
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


def remap(a, pe1, pe2, iv=0, qs=None, dp2=None):
    dp1 = pe1[:, 1:] - pe1[:, :-1]
    qi = cs_profile(a, dp1, iv, qs)
    return remap_integrate(a, qi, pe1, pe2, dp2)


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
    pt = remap(pt, pn1, pn2, iv=1)
    qn = [remap(q[C], pe1, pe2, iv=0, dp2=dp2) for q in st["q"]]
    out = dict(st)
    if not hydro:
        w = remap(st["w"][C], pe1, pe2, iv=-2, qs=st["ws"][C][:, 0])
        delz = remap(delz, pe1, pe2, iv=1)
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
    u = remap(st["u"][Cu], pe0, pe3, iv=-1)
    Cv = (slice(None), slice(None), R(js, je), R(is_, ie + 1))
    Cvm = (slice(None), slice(None), R(js, je), R(is_ - 1, ie))
    pe0 = 0.5 * (pe[Cvm] + pe[Cv])
    pe0 = torch.cat([pe[Cv][:, :1] * 0 + pe[Cv][:, :1], pe0[:, 1:]], dim=1)
    pe3 = AK + 0.5 * BK * (pe[Cvm][:, K:] + pe[Cv][:, K:])
    pe3 = torch.cat([0 * pe3[:, :1] + float(ak[0]), pe3[:, 1:]], dim=1)
    v = remap(st["v"][Cv], pe0, pe3, iv=-1)
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
