"""BASELINE config 2: C48 L72 dynamics-only TL/AD window on one B200 -- a 3-step window with the trajectory
resident on the device (step_nl fills the slots), TL forward and AD backward through the window, the window-level
dot-product test and a Taylor test of the TL against finite differences of the nonlinear window.  Size-independent
properties only (the oracle cannot run C48 L72 in seconds)."""
import numpy as np
import pytest
import fv3lm
from synth import grid as G, state as S

ZVIR = (8314.47 / 18.015) / (8314.47 / 28.965) - 1.0


def _window(nonhydro, N=48, K=72, nsteps=3):
    dt = 450.0 * 180.0 / N
    ak, bk = S.eta_levels(K)
    M = G.build_metrics(N)
    cfg = fv3lm.default_config(N, K, n_split=S.n_split_auto(N, dt, not nonhydro), k_split=1, dt=dt, ptop=1.0, d2_bg_k1=0.20, d2_bg_k2=0.10,
                               hydrostatic=0 if nonhydro else 1, zvir=ZVIR)
    h = fv3lm.FV3LM(cfg, ak, bk)
    h.set_metrics(M)
    st = S.make_state(M, K, ak, bk, hydrostatic=not nonhydro)
    fields = [f for f in h.FIELDS if f in st]
    h.set_phis(st["phis"])
    h.traj_set(0, {k: st[k] for k in fields})
    for n in range(nsteps):                       # trajectory window stays on the device
        h.step_nl(n, n + 1)
    base = {k: np.zeros_like(st[k]) for k in fields}
    h.traj_get(nsteps, base)
    assert all(np.isfinite(base[k]).all() for k in fields)
    dx = S.make_pert(st, 7)
    y = S.make_pert(st, 8)
    # TL through the window, increments resident on the device
    h.pert_upload({k: dx[k] for k in fields})
    for n in range(nsteps):
        h.step_tl_dev(n)
    mdx = {k: np.zeros_like(st[k]) for k in fields}
    h.pert_download(mdx)
    # AD back through the window
    h.pert_upload({k: y[k] for k in fields})
    for n in reversed(range(nsteps)):
        h.step_ad_dev(n)
    mty = {k: np.zeros_like(st[k]) for k in fields}
    h.pert_download(mty)
    lhs = sum((mdx[k] * y[k]).sum() for k in fields)
    rhs = sum((dx[k] * mty[k]).sum() for k in fields)
    dot = abs(lhs - rhs) / max(abs(lhs), abs(rhs))
    assert dot <= 1e-10, (lhs, rhs, dot)
    # Taylor test of the window
    ratios = []
    for eps in (1e-2, 1e-3, 1e-4):    # (the increments are O(10 Pa) in delp: larger steps leave the thin top layers)
        h.traj_set(nsteps + 1, {k: st[k] + eps * dx[k] for k in fields})
        for n in range(nsteps):
            h.step_nl(nsteps + 1 + n, nsteps + 2 + n)
        outp = {k: np.zeros_like(st[k]) for k in fields}
        h.traj_get(2 * nsteps + 1, outp)
        num = den = 0.0
        for k in fields:
            sc = 1.0 / (np.abs(mdx[k]).max() + 1e-300)
            num += ((((outp[k] - base[k]) - eps * mdx[k]) * sc) ** 2).sum()
            den += ((eps * mdx[k] * sc) ** 2).sum()
        ratios.append(float(np.sqrt(num / den)))
    assert ratios[1] < 0.25 * ratios[0] and ratios[2] < 0.25 * ratios[1], ratios
    h.close()
    return dict(dot=dot, taylor=ratios)


@pytest.mark.gpu
@pytest.mark.parametrize("nonhydro", [False, True])
def test_c48_l72_window_gpu(nonhydro):
    print(_window(nonhydro))
