"""Synthetic, deterministic model state for tests and bench.py (SURVEY 8(d) "synthetic
inputs"): analytic hybrid sigma-p levels, a zonal jet + seeded large-scale noise on the
D grid, a Gaussian mountain, a lapse-rate temperature profile and smooth tracers.
Everything is fp64 and generated on the compute domain [6, K, N, N] (Fortran (i,j,k) per tile).
"""
import numpy as np
from .cubed_sphere import R
from . import grid as G

RDGAS = 8314.47 / 28.965
GRAV = 9.80665


def eta_levels(K, ptop=1.0, p0=1.0e5, pc=1.5e4):
    """ak, bk (K+1) of an analytic hybrid coordinate: pure pressure above pc, ak(K)=0, bk(K)=1"""
    s = np.linspace(0.0, 1.0, K + 1)
    pref = ptop * np.exp(np.log(p0 / ptop) * s ** 0.85)
    pref[0] = ptop; pref[-1] = p0
    bk = np.clip((pref - pc) / (p0 - pc), 0.0, 1.0) ** 1.3
    bk[0] = 0.0; bk[-1] = 1.0
    ak = pref - bk * p0
    ak[-1] = 0.0; ak[0] = ptop
    return ak, bk


def n_split_auto(N, dt, hydrostatic, k_split=1):
    """model/fv_control_nlm.F90:709-749 (ns0 = 5, dim0 = 180, dt0 = 1800 defaults)"""
    npx = N + 1
    ns0 = 5
    if hydrostatic:
        if npx >= 120:
            ns0 = 6
    else:
        ns0 = 6 if npx <= 45 else (7 if npx <= 90 else 8)
    n0 = max(1, int(np.floor(ns0 * abs(dt) * 4.0 * N / (1800.0 * 180.0) + 0.49 + 0.5)))
    return int(np.floor(n0 / float(k_split) + 0.5 + 0.5))


def _smooth_noise(rng, shape_lo, N):
    """band-limited noise: coarse random field interpolated bilinearly to N x N"""
    lo = rng.standard_normal(shape_lo)
    n = shape_lo[-1]
    x = np.linspace(0.0, n - 1.0, N)
    i0 = np.clip(np.floor(x).astype(int), 0, n - 2); w = x - i0
    a = lo[..., :, i0] * (1 - w) + lo[..., :, i0 + 1] * w
    b = a[..., i0, :] * (1 - w)[:, None] + a[..., i0 + 1, :] * w[:, None]
    return b


def make_state(M, K, ak, bk, seed=20261018, hydrostatic=True, u0=30.0):
    """M: metrics dict (synth.grid.build_metrics).  Returns dict of compute-domain arrays
    u v t delp qv ql qi o3 [w delz] ([6,K,N,N]) and phis ([6,N,N])."""
    N = M["N"]
    rng = np.random.default_rng(seed)
    grid = M["grid"]; agrid = M["agrid"]
    c = (slice(None), R(1, N), R(1, N))
    lon = agrid[c + (0,)]; lat = agrid[c + (1,)]
    # Gaussian mountain and surface pressure
    d = G.gc_dist(np.stack([lon, lat], -1), np.array([np.pi / 2, np.pi / 6]))
    zs = 1500.0 * np.exp(-(d / 0.3) ** 2)
    phis = GRAV * zs
    T0 = 288.0
    ps = 1.0e5 * np.exp(-phis / (RDGAS * T0)) + 200.0 * _smooth_noise(rng, (6, 7, 7), N)
    pe = ak[None, :, None, None] + bk[None, :, None, None] * ps[:, None]
    delp = pe[:, 1:] - pe[:, :-1]
    pm = 0.5 * (pe[:, 1:] + pe[:, :-1])
    t = np.maximum(T0 * (pm / 1.0e5) ** 0.19, 210.0) + 1.5 * _smooth_noise(rng, (6, K, 7, 7), N)
    # winds: zonal jet u0 cos(lat) + noise, projected on the D-grid edge directions
    xyz = G.ll2xyz(grid)
    def edge_wind(p_a, p_b):
        mid = G.mid3(p_a, p_b)
        e = p_b - p_a; e = e / np.linalg.norm(e, axis=-1, keepdims=True)
        ll = G.xyz2ll(mid)
        east = np.stack([-np.sin(ll[..., 0]), np.cos(ll[..., 0]), np.zeros_like(ll[..., 0])], -1)
        return (east * e).sum(-1) * np.cos(ll[..., 1])
    cu = edge_wind(xyz[:, R(1, N), R(1, N)], xyz[:, R(1, N), R(2, N + 1)])      # u(i,j): edge (i,j)-(i+1,j)
    cv = edge_wind(xyz[:, R(1, N), R(1, N)], xyz[:, R(2, N + 1), R(1, N)])      # v(i,j): edge (i,j)-(i,j+1)
    prof = (u0 * np.sin(np.pi * np.clip(pm.mean(axis=(0, 2, 3)) / 1.0e5, 0, 1)) ** 0.5)[None, :, None, None]
    u = prof * cu[:, None] + 2.0 * _smooth_noise(rng, (6, K, 9, 9), N)
    v = prof * cv[:, None] + 2.0 * _smooth_noise(rng, (6, K, 9, 9), N)
    qv = 0.015 * (pm / 1.0e5) ** 3 * (1.0 + 0.2 * _smooth_noise(rng, (6, K, 7, 7), N))
    ql = 1.0e-6 * (1.0 + 0.3 * _smooth_noise(rng, (6, K, 7, 7), N))
    qi = 1.0e-6 * (1.0 + 0.3 * _smooth_noise(rng, (6, K, 7, 7), N))
    o3 = 1.0e-6 * (1.0 + 0.3 * _smooth_noise(rng, (6, K, 7, 7), N))
    st = dict(u=u, v=v, t=t, delp=delp, qv=qv, ql=ql, qi=qi, o3=o3, phis=phis)
    if not hydrostatic:
        peln = np.log(pe)
        st["w"] = 0.05 * _smooth_noise(rng, (6, K, 7, 7), N)
        st["delz"] = -(RDGAS * t * (1.0 + 0.6078 * qv) / GRAV) * (peln[:, 1:] - peln[:, :-1])
    return {k: np.ascontiguousarray(a, dtype=np.float64) for k, a in st.items()}


def make_pert(st, seed, scale=1.0):
    """seeded perturbation / adjoint test vector with field-appropriate magnitudes"""
    rng = np.random.default_rng(seed)
    mag = dict(u=1.0, v=1.0, t=1.0, delp=10.0, qv=1e-4, ql=1e-7, qi=1e-7, o3=1e-8, w=0.1, delz=1.0)
    out = {}
    for k, a in st.items():
        if k == "phis":
            continue
        N = a.shape[-1]
        out[k] = np.ascontiguousarray(scale * mag[k] * _smooth_noise(rng, a.shape[:2] + (max(7, N // 6), max(7, N // 6)), N))
    return out
