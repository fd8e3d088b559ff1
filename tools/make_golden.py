#!/usr/bin/env python
"""Regenerates tests/golden/*.npz: outputs of the ORACLE (torch fp64 restatement) for one C12 fv3jedi_lm dynamics
step, hydrostatic, non-hydrostatic and non-hydrostatic in two-sided mode (both flag structures at the reference's defaults): NL result, TL result (jvp) for a seeded increment and AD result (vjp) for a
seeded adjoint vector; physics_c12.npz: turbulence TL / AD, cubed_to_latlon, tracer_2d with q_split = 0.  The reference itself cannot be run here (no Fortran/FMS/MPI), so these fixtures pin the
oracle against silent regressions between rounds -- they are NOT reference outputs (the reference-executed fixtures are tests/golden/ref_*.npz, written by tests/golden/make_ref_golden.py).
  python tools/make_golden.py
"""
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "fv3-jedi-linearmodel_b200")):
    sys.path.insert(0, p)


def case(nonhydro, two_sided=False):
    from test_multirank import _inputs
    from oracle import fv_dynamics as ofv
    from common import ograd
    N, K, ak, bk, f, act, p, dx, y = _inputs(nonhydro, two_sided)
    g = ograd(N)
    cfg = dict(p["_oracle_cfg"]) if two_sided else dict(p)
    cfg.update(hydrostatic=not nonhydro, nq=4, bdt=900.0)
    phis = torch.from_numpy(f["phis"])
    def fn(*a):
        o = ofv.step_nl(dict(zip(act, a)), g, ak, bk, cfg, phis)
        return tuple(o[k] for k in act)
    x = tuple(torch.from_numpy(f[k]) for k in act)
    nl, tl = torch.func.jvp(fn, x, tuple(torch.from_numpy(dx[k]) for k in act))
    _, vjp = torch.func.vjp(fn, *x)
    ad = vjp(tuple(torch.from_numpy(y[k + "_n"]) for k in act))
    out = {}
    o = 2
    for k, a, b, c in zip(act, nl, tl, ad):
        out["nl." + k] = a.numpy()[:, :, o + 1:o + 1 + N, o + 1:o + 1 + N]
        out["tl." + k] = b.numpy()[:, :, o + 1:o + 1 + N, o + 1:o + 1 + N]
        out["ad." + k] = c.numpy()[:, :, o + 1:o + 1 + N, o + 1:o + 1 + N]
    return out


def case_physics():
    """the pieces added around the dynamics step: linearised turbulence (TL, AD), cubed_to_latlon, tracer_2d with q_split = 0"""
    import test_turbulence as tt
    import test_tracer_2d as ttr
    from test_dyn_core import CFG
    from oracle import turbulence as otb, c2l as oc, fv_dynamics as ofv
    from oracle.dyn_core import halo_of
    from common import ograd, rnd
    out = {}
    shape, co, traj, dx, y = tt.golden_inputs()
    lt = otb.set_ltraj(co, traj["delp"], CFG["ptop"], tt.KAPPA)
    tl = otb.step(lt, dx, tt.KAPPA, 1); ad = otb.step(lt, y, tt.KAPPA, 2)
    for n in tt.FLD:
        out["turb.tl." + n] = tl[n]; out["turb.ad." + n] = ad[n]
    N, K = 12, 2
    halo, _ = halo_of(N)
    rng = np.random.default_rng(5)
    u, v = 10.0 * rnd(rng, N, K), 10.0 * rnd(rng, N, K)
    uh, vh = halo.dgrid(torch.from_numpy(u), torch.from_numpy(v))
    ua, va = oc.c2l_ord4(uh, vh, ograd(N))
    o = 2
    out["c2l.ua"] = ua.numpy()[:, :, o + 1:o + 1 + N, o + 1:o + 1 + N]; out["c2l.va"] = va.numpy()[:, :, o + 1:o + 1 + N, o + 1:o + 1 + N]
    f, _ = ttr.fields(N, 6, ttr.AMP3, 77)
    t = {k: torch.from_numpy(a) for k, a in f.items()}
    q = ofv.tracer_2d([halo.scalar(t["q0"]), halo.scalar(t["q1"])], t["dp1"], t["mfx"], t["mfy"], t["cx"], t["cy"], ograd(N), 2,
                      q_split=0, q_split_max=3, halo=halo)
    for n, a in zip(("q0", "q1"), q):
        out["trc." + n] = a.numpy()[:, :, o + 1:o + 1 + N, o + 1:o + 1 + N]
    return out


if __name__ == "__main__":
    torch.set_default_dtype(torch.float64)
    if len(sys.argv) < 2 or sys.argv[1] == "physics":
        path = os.path.join(ROOT, "tests", "golden", "physics_c12.npz")
        np.savez_compressed(path, **case_physics())
        print(path, os.path.getsize(path))
    for nh, two in ((False, False), (True, False), (True, True)):
        if len(sys.argv) > 1 and sys.argv[1] == "physics":
            break
        if len(sys.argv) > 1 and sys.argv[1] == "two_sided" and not two:
            continue                        # (regenerate only the new fixture)
        out = case(nh, two)
        path = os.path.join(ROOT, "tests", "golden", "step_c12_%s%s.npz" % ("nonhydro" if nh else "hydro", "_two_sided" if two else ""))
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path))
