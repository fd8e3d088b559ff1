"""c_sw (C-grid half step): C-ABI library vs the torch oracle."""
import numpy as np
import pytest
import torch
from oracle import sw_core as osw
from common import metrics, ograd, handle, rnd, check_module


def smooth_state(N, K, seed, nonhydro=True):
    rng = np.random.default_rng(seed)
    f = {}
    f["delp"] = 1000.0 + 50.0 * rnd(rng, N, K)
    f["pt"] = 300.0 + 5.0 * rnd(rng, N, K)
    f["u"] = 10.0 * rnd(rng, N, K)
    f["v"] = 10.0 * rnd(rng, N, K)
    f["w"] = 0.5 * rnd(rng, N, K)
    return f, rng


def _run(emu, hydrostatic):
    N, K = 12, 2
    f, rng = smooth_state(N, K, 7)
    g = ograd(N)
    dt2 = 225.0
    names = ["delp", "pt", "u", "v", "w"]
    act = names if not hydrostatic else ["delp", "pt", "u", "v"]
    onames = ["delpc", "ptc", "uc", "vc", "ua", "va", "ut", "vt", "divg_d"] + ([] if hydrostatic else ["wc"])
    def fn(*a):
        d = dict(zip(act, a))
        w = d.get("w", torch.from_numpy(f["w"]))
        o = osw.c_sw(d["delp"], d["pt"], d["u"], d["v"], w, g, dt2, hydrostatic, 1)
        return tuple(o[k] for k in onames)
    npx = N + 1
    outs = {"delpc": (0, npx, 0, npx), "ptc": (0, npx, 0, npx), "uc": (1, npx, 1, N), "vc": (1, N, 1, npx),
            "ua": (0, npx, 0, npx), "va": (0, npx, 0, npx), "ut": (0, npx + 1, 0, npx), "vt": (0, npx, 0, npx + 1),
            "divg_d": (1, npx, 1, npx)}
    if not hydrostatic:
        outs["wc"] = (0, npx, 0, npx)
    outs = {k: outs[k] for k in onames}
    h = handle(N, K, emu)
    res = check_module(h, "c_sw", N, K, f, act, outs, fn, {"dt2": dt2, "hydrostatic": int(hydrostatic), "nord": 1}, rng)
    return res


@pytest.mark.parametrize("hydrostatic", [True, False])
def test_c_sw_emu(hydrostatic):
    _run(True, hydrostatic)


@pytest.mark.gpu
@pytest.mark.parametrize("hydrostatic", [True, False])
def test_c_sw_gpu(hydrostatic):
    _run(False, hydrostatic)
