"""A compiled C program (tests/c_driver.c, gcc) drives the C ABI end to end -- create, set_metric*, set_phis, traj_set, step_nl,
traj_get, step_tl, step_ad, destroy -- and must return bit-for-bit what the ctypes binding returns for the same inputs.  It is the
closest stand-in for the Fortran ISO_C_BINDING shim (fortran/fv3jedi_lm_dynamics_mod.F90), which this image cannot compile."""
import ctypes
import os
import struct
import subprocess
import tempfile
import numpy as np
import pytest
import fv3lm
from common import metrics
from synth import state as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ZVIR = (8314.47 / 18.015) / (8314.47 / 28.965) - 1.0


def build_driver(emu):
    lib = fv3lm.lib_path(emu)
    exe = os.path.join(tempfile.gettempdir(), "fv3lm_c_driver_%s_%d" % ("emu" if emu else "cuda", os.getuid()))
    src = os.path.join(ROOT, "tests", "c_driver.c")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(lib)):
        subprocess.check_call(["gcc", "-O1", "-std=c99", "-o", exe, src, lib, "-Wl,-rpath," + os.path.dirname(lib), "-lm"])
    return exe


def rec(fh, name, kind, a):
    a = np.ascontiguousarray(a, dtype=np.float64).ravel()
    fh.write(name.encode().ljust(32, b"\0")); fh.write(struct.pack("<iq", kind, a.size)); fh.write(a.tobytes())


def _run(emu, nonhydro=True):
    N, K = 12, 6
    ak, bk = S.eta_levels(K)
    M = metrics(N)
    st = S.make_state(M, K, ak, bk, hydrostatic=not nonhydro)
    fields = [f for f in fv3lm.FV3LM.FIELDS if f in st]
    pert = S.make_pert(st, 5); yvec = S.make_pert(st, 6)
    kw = dict(n_split=2, k_split=1, dt=1800.0, ptop=1.0, d2_bg_k1=0.2, d2_bg_k2=0.1, hydrostatic=0 if nonhydro else 1, zvir=ZVIR)
    cfg = fv3lm.default_config(N, K, **kw)
    # ---- the ctypes path
    h = fv3lm.FV3LM(cfg, ak, bk, emu=emu)
    h.set_metrics(M)
    h.set_phis(st["phis"])
    h.traj_set(0, {k: st[k] for k in fields})
    h.step_nl(0, 1)
    ref = {}
    out = {k: np.zeros_like(st[k]) for k in fields}
    h.traj_get(1, out)
    for k in fields:
        ref["nl." + k] = out[k]
    a = {k: pert[k].copy() for k in fields}; h.step_tl(0, a)
    b = {k: yvec[k].copy() for k in fields}; h.step_ad(0, b)
    for k in fields:
        ref["tl." + k] = a[k]; ref["ad." + k] = b[k]
    # ---- the same inputs through the compiled driver
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
        with open(fin, "wb") as fh:
            raw = bytes(memoryview(cfg))
            fh.write(b"FV3LMIN1"); fh.write(struct.pack("<i", len(raw))); fh.write(raw)
            rec(fh, "ak", 3, ak); rec(fh, "bk", 4, bk)
            for n in fv3lm.METRICS_2D:
                rec(fh, n, 0, M[n])
            for k in (1, 2, 3, 4):
                rec(fh, "sin_sg%d" % k, 0, M["sin_sg"][..., k]); rec(fh, "cos_sg%d" % k, 0, M["cos_sg"][..., k])
            rec(fh, "agrid_lon", 0, M["agrid"][..., 0]); rec(fh, "agrid_lat", 0, M["agrid"][..., 1])
            rec(fh, "grid_lon", 0, M["grid"][..., 0]); rec(fh, "grid_lat", 0, M["grid"][..., 1])
            for n in fv3lm.METRICS_1D:
                rec(fh, n, 1, M[n])
            for n in ("da_min", "da_min_c"):
                rec(fh, n, 2, [M[n]])
            rec(fh, "phis", 5, st["phis"])
            for k in fields:
                rec(fh, k, 6, st[k]); rec(fh, k, 7, pert[k]); rec(fh, k, 8, yvec[k])
        p = subprocess.run([build_driver(emu), fin, fout], capture_output=True, text=True, timeout=900)
        assert p.returncode == 0, p.stdout + p.stderr
        got = {}
        with open(fout, "rb") as fh:
            while True:
                nm = fh.read(32)
                if len(nm) < 32:
                    break
                n, = struct.unpack("<q", fh.read(8))
                got[nm.rstrip(b"\0").decode()] = np.frombuffer(fh.read(8 * n), dtype=np.float64)
    assert sorted(got) == sorted(ref)
    for k in ref:
        assert np.array_equal(got[k], ref[k].ravel()), k        # same library, same inputs, same call sequence: bit-identical
    return p.stdout.strip()


def test_c_driver_emu():
    print(_run(True))


def test_c_driver_hydro_emu():
    print(_run(True, nonhydro=False))


@pytest.mark.gpu
def test_c_driver_gpu():
    print(_run(False))
