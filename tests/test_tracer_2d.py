"""tracer_2d with q_split = 0 (model/fv_tracer2d_nlm.F90:351-502, TL model_tlmadm/fv_tracer2d_tlm.F90:885-1100): the number of
tracer sub-steps follows the accumulated Courant numbers, per level.  C ABI vs the torch oracle for NL, TL (jvp), AD (vjp) and
the dot-product identity, on Courant numbers that put different levels on 1, 2 and 3 sub-steps."""
import numpy as np
import pytest
import torch
from oracle import fv_dynamics as ofv
from oracle.dyn_core import halo_of
from common import ograd, handle, check_module, metrics
from test_dyn_core import CFG, smooth

# per-level Courant amplitude -> (with the 1 - sin_sg(5) <= 0.1 term of levels k >= npz/6) ksplt = 1 1 2 3 1 2 ...
AMP3 = [0.93, 0.45, 1.35, 2.25, 0.30, 1.60, 0.70, 2.40, 0.10, 1.20, 0.55, 0.80]
AMP1 = [0.60, 0.45, 0.35, 0.25, 0.30, 0.60, 0.70, 0.40, 0.10, 0.20, 0.55, 0.80]


def fields(N, K, amp, seed):
    rng = np.random.default_rng(seed)
    m = metrics(N)
    a = np.asarray(amp[:K])[None, :, None, None]
    shape = lambda rng: 0.97 + 0.03 * np.tanh(smooth(rng, N, K))          # in (0.94, 1]: max |c| of level k lies in (0.94, 1] amp(k)
    sx = np.where(smooth(rng, N, K) > 0.0, 1.0, -1.0)                     # both upwind branches of xfx / yfx
    f = dict(q0=0.01 * (1.0 + 0.3 * smooth(rng, N, K)), q1=1e-5 * (1.0 + 0.3 * smooth(rng, N, K)),
             dp1=1000.0 * (1.0 + 0.05 * smooth(rng, N, K)))
    cx = a * shape(rng) * sx
    cy = -a * shape(rng) * sx
    area = np.asarray(m["area"])[:, None]
    f["mfx"] = cx * area * 30.0 * (1.0 + 0.1 * smooth(rng, N, K))          # |div| / area << dp1: dp2 stays positive
    f["mfy"] = cy * area * 30.0 * (1.0 + 0.1 * smooth(rng, N, K))
    f["cx"] = cx; f["cy"] = cy
    return f, rng


def _run(emu, amp=AMP3, K=12, N=12, q_split_max=3, hord=2, modes=("nl", "tl", "ad"), expect_nsplt=3, extra=None, seed=77):
    f, rng = fields(N, K, amp, seed)
    g = ograd(N)
    halo, _ = halo_of(N)
    act = list(f.keys())
    hord_t = (extra or {}).get("t.hord_tr")
    def fn(*a):
        d = dict(zip(act, a))
        q = [halo.scalar(d["q0"]), halo.scalar(d["q1"])]
        o = ofv.tracer_2d(q, d["dp1"], d["mfx"], d["mfy"], d["cx"], d["cy"], g, hord_t if hord_t else hord, hord if hord_t else None,
                          q_split=0, q_split_max=q_split_max, halo=halo)
        return tuple(o)
    C = (1, N, 1, N)
    outs = dict(q0_n=C, q1_n=C)
    h = handle(N, K, emu)
    p = dict(CFG); p.update(hord_tr=hord, q_split=0, q_split_max=q_split_max)
    if extra:
        p.update(extra)
    r = check_module(h, "tracer_2d", N, K, f, act, outs, fn, p, rng, tol=1e-11, dot_tol=1e-12, pert_scale=1e-3, modes=modes)
    if expect_nsplt is not None:
        assert ofv.tracer_2d.last_nsplt == expect_nsplt, ofv.tracer_2d.last_nsplt
    return r


def test_sub_steps_emu():
    """levels on 1, 2 and 3 sub-steps (nsplt = 3)"""
    print(_run(True))


def test_sub_steps_spare_emu():
    """q_split_max = 4 > nsplt = 3: the spare sub-step passes every level through"""
    print(_run(True, q_split_max=4, K=8))


def test_single_step_emu():
    """all Courant numbers below 1: nsplt = 1, nothing is scaled (the q_split = 1 result)"""
    r = _run(True, amp=AMP1, K=7, expect_nsplt=1)
    print(r)
    f, rng = fields(12, 7, AMP1, 77)
    g = ograd(12); halo, _ = halo_of(12)
    t = {k: torch.from_numpy(v) for k, v in f.items()}
    q = [halo.scalar(t["q0"]), halo.scalar(t["q1"])]
    a = ofv.tracer_2d(q, t["dp1"], t["mfx"], t["mfy"], t["cx"], t["cy"], g, 2, q_split=0, halo=halo)
    b = ofv.tracer_2d(q, t["dp1"], t["mfx"], t["mfy"], t["cx"], t["cy"], g, 2, q_split=1)
    from oracle.cubed_sphere import R
    assert all(torch.equal(x[..., R(1, 12), R(1, 12)], y[..., R(1, 12), R(1, 12)]) for x, y in zip(a, b))


def test_two_sided_emu():
    """split_hord: monotone scheme for the trajectory, linear scheme for the increment, in every sub-step"""
    print(_run(True, K=6, extra={"two_sided": 1, "t.hord_tr": 8}))


def test_too_many_sub_steps_emu():
    """nsplt = 3 > q_split_max = 2: the run-time check turns into an error of the call"""
    with pytest.raises(RuntimeError, match="q_split_max"):
        _run(True, q_split_max=2, K=6, modes=("nl",))


@pytest.mark.gpu
def test_sub_steps_gpu():
    _run(False)


@pytest.mark.gpu
def test_two_sided_gpu():
    _run(False, K=6, extra={"two_sided": 1, "t.hord_tr": 8})


@pytest.mark.gpu
def test_too_many_sub_steps_gpu():
    with pytest.raises(RuntimeError, match="q_split_max"):
        _run(False, q_split_max=2, K=6, modes=("nl",))
