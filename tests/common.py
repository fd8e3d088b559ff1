"""Shared helpers for the parity tests: build a small cubed sphere, seeded fields, and
run a module through the C ABI (GPU library, or the test-only host emulation)."""
import functools
import numpy as np
import torch
import fv3lm
from oracle import grid as ogrid
from oracle.tp_core import Grid
from oracle.cubed_sphere import R

torch.set_default_dtype(torch.float64)


@functools.lru_cache(maxsize=4)
def metrics(N):
    return ogrid.build_metrics(N)


@functools.lru_cache(maxsize=4)
def ograd(N):
    return Grid(metrics(N))


_handles = {}


def handle(N, K, emu, **kw):
    key = (N, K, emu, tuple(sorted(kw.items())))
    if key not in _handles:
        cfg = fv3lm.default_config(N, K, **kw)
        h = fv3lm.FV3LM(cfg, emu=emu)
        h.set_metrics(metrics(N))
        _handles[key] = h
    return _handles[key]


def rnd(rng, N, K, scale=1.0, mean=0.0):
    NX = N + 7
    return mean + scale * rng.standard_normal((6, K, NX, NX))


def relerr(a, b):
    a = np.asarray(a); b = np.asarray(b)
    den = max(np.abs(b).max(), 1e-300)
    return float(np.abs(a - b).max() / den)


def region(a, i0, i1, j0, j1):
    return a[..., R(j0, j1), R(i0, i1)]
