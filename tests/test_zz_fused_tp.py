"""Routine-level kernels (default since round 2; the file keeps its round-1 name, it used to run last because the kernels had not been on a
B200 yet): fv_tp_2d as shared-memory-tile kernels (csrc/fused_tp.h), the row-marching variant, a2b_ord4 as tile kernel and as compiled operator,
automatically fused stage chains.  Forward sweeps (NL, TL) of the tile kernels (FV3LM_FUSED_TP=1) must reproduce the
stage-by-stage chain of model_tlmadm/tp_core_tlm.F90:2123-2324 -- same arithmetic, so the comparison is to round-off -- and the
adjoint (which keeps the chain) must stay the transpose of the fused tangent."""
import numpy as np
import pytest
import fv3lm
import common
from common import metrics, handle, rnd, relerr, region
import test_tp_core
import test_step_api
import test_d_sw
import test_decomp
import test_c_sw

TOL = 1e-13     # identical expressions; the device build may contract multiply-adds differently in the two kernels


@pytest.fixture
def fused2(monkeypatch):
    """FV3LM_FUSED_TP=2: the adjoint runs tile kernels as well (no stage chain in the program)"""
    common._handles.clear()
    monkeypatch.setenv("FV3LM_FUSED_TP", "2")
    yield
    common._handles.clear()


@pytest.fixture
def fused(monkeypatch):
    """programs are built once per handle and read the switch at build time: start from (and leave) an empty handle cache"""
    common._handles.clear()
    monkeypatch.setenv("FV3LM_FUSED_TP", "1")
    yield
    common._handles.clear()


def _fields(N, K, seed):
    f, rng = test_tp_core._inputs(N, K, seed)
    f["mfx"] = rnd(rng, N, K, 1.0) * np.abs(metrics(N)["area"][:, None])
    f["mfy"] = rnd(rng, N, K, 1.0) * np.abs(metrics(N)["area"][:, None])
    return f, rng


def _module(emu, N, K, f, dp, mode, params):
    h = handle(N, K, emu)
    names = [n for n in f if n not in ("mfx", "mfy") or params.get("use_mf")]
    traj = {n: f[n].copy() for n in names}
    traj.update(fx=np.zeros_like(f["q"]), fy=np.zeros_like(f["q"]))
    pert = None
    if mode == fv3lm.MODE_TL:
        pert = {n: dp[n].copy() for n in names}
        pert.update(fx=np.zeros_like(f["q"]), fy=np.zeros_like(f["q"]))
    h.module_run("fv_tp_2d", mode, traj, pert, params=params)
    return traj, pert


def _fused_vs_chain(emu, monkeypatch, N, K, params, tl=True):
    f, rng = _fields(N, K, 77)
    dp = {n: rnd(rng, N, K) * (np.abs(f[n]).mean() * 1e-2) for n in f}
    res = {}
    for flag in ("0", "1"):
        common._handles.clear()
        monkeypatch.setenv("FV3LM_FUSED_TP", flag)
        res[flag] = [_module(emu, N, K, f, dp, fv3lm.MODE_NL, params)]
        if tl:
            res[flag].append(_module(emu, N, K, f, dp, fv3lm.MODE_TL, params))
    common._handles.clear()
    for a, b in zip(res["0"], res["1"]):
        for which in (0, 1):
            if a[which] is None:
                continue
            fx0, fx1 = region(a[which]["fx"], 1, N + 1, 1, N), region(b[which]["fx"], 1, N + 1, 1, N)
            fy0, fy1 = region(a[which]["fy"], 1, N, 1, N + 1), region(b[which]["fy"], 1, N, 1, N + 1)
            assert np.abs(fx0).max() > 0 and np.abs(fy0).max() > 0
            assert relerr(fx1, fx0) <= TOL and relerr(fy1, fy0) <= TOL, (params, which, relerr(fx1, fx0), relerr(fy1, fy0))


def _module_ad(emu, N, K, f, y, params):
    h = handle(N, K, emu)
    names = [n for n in f if n not in ("mfx", "mfy") or params.get("use_mf")]
    traj = {n: f[n].copy() for n in names}
    traj.update(fx=np.zeros_like(f["q"]), fy=np.zeros_like(f["q"]))
    pert = {n: np.zeros_like(f[n]) for n in names}
    pert.update(fx=y["fx"].copy(), fy=y["fy"].copy())
    h.module_run("fv_tp_2d", fv3lm.MODE_AD, traj, pert, params=params)
    return {n: pert[n] for n in names}


def _rev_vs_chain(emu, monkeypatch, N, K, params):
    """reverse tile kernels (level 2) against the stage chain's gather adjoints, all input adjoints"""
    f, rng = _fields(N, K, 78)
    y = dict(fx=np.zeros_like(f["q"]), fy=np.zeros_like(f["q"]))
    region(y["fx"], 1, N + 1, 1, N)[...] = region(rnd(rng, N, K), 1, N + 1, 1, N)
    region(y["fy"], 1, N, 1, N + 1)[...] = region(rnd(rng, N, K), 1, N, 1, N + 1)
    res = {}
    for flag in ("0", "2"):
        common._handles.clear()
        monkeypatch.setenv("FV3LM_FUSED_TP", flag)
        res[flag] = _module_ad(emu, N, K, f, y, params)
    common._handles.clear()
    for n in res["0"]:
        assert np.abs(res["0"][n]).max() > 0 or (params["hord"] == 1 and n in ("crx", "cry")), n     # (upwind fluxes do not depend on c)
        assert relerr(res["2"][n], res["0"][n]) <= 1e-12, (params, n, relerr(res["2"][n], res["0"][n]))


@pytest.mark.parametrize("params", [dict(hord=2), dict(hord=1), dict(hord=333), dict(hord=2, use_mf=1), dict(hord=2, n_sponge=1)])
def test_reverse_vs_chain_emu(monkeypatch, params):
    _rev_vs_chain(True, monkeypatch, 40, 2, params)


@pytest.mark.parametrize("hord", [1, 2, 333])
def test_fused2_vs_oracle_emu(fused2, hord):
    """NL / TL / AD all through tile kernels against the oracle (jvp / vjp) + dot-product test"""
    test_tp_core._run(True, hord)


def test_fused2_step_emu(fused2):
    print(test_step_api._run(True, nonhydro=True))


def test_fused2_step_hydro_segmented_emu(fused2, monkeypatch):
    """hydrostatic step with the checkpoint / recompute adjoint forced"""
    monkeypatch.setenv("FV3LM_AD_STORE_BUDGET", "0")
    print(test_step_api._run(True, nonhydro=False))


def test_fused2_step_two_sided_emu(fused2):
    test_step_api.test_step_api_reference_defaults_emu()


@pytest.mark.gpu
@pytest.mark.parametrize("params", [dict(hord=2), dict(hord=333), dict(hord=2, use_mf=1), dict(hord=1, n_sponge=1)])
def test_reverse_vs_chain_gpu(monkeypatch, params):
    _rev_vs_chain(False, monkeypatch, 40, 2, params)


@pytest.mark.gpu
def test_fused2_step_gpu(fused2):
    print(test_step_api._run(False, nonhydro=True))


# N = 40: the padded array (47 x 47) spans two tiles in x and six in y, so footprints cross block boundaries on both axes
@pytest.mark.parametrize("params", [dict(hord=2), dict(hord=1), dict(hord=333), dict(hord=2, use_mf=1), dict(hord=2, n_sponge=1)])
def test_fused_vs_chain_emu(monkeypatch, params):
    _fused_vs_chain(True, monkeypatch, 40, 2, params)


@pytest.mark.parametrize("hord", [5, 8, 9, 13])
def test_fused_vs_chain_trajectory_schemes_emu(monkeypatch, hord):
    _fused_vs_chain(True, monkeypatch, 40, 2, dict(hord=hord), tl=False)


@pytest.mark.parametrize("hord", [1, 2, 333])
def test_fused_vs_oracle_emu(fused, hord):
    """NL / TL through the fused kernels and AD through the chain against the oracle, dot-product test across the two"""
    test_tp_core._run(True, hord)


def test_fused_step_emu(fused):
    """a whole model step: step_nl / step_tl run the fused transports, step_ad the chain -- oracle parity, dot-product and Taylor tests"""
    print(test_step_api._run(True, nonhydro=True))


def test_fused_step_two_sided_emu(fused):
    """two-sided configuration: the trajectory-side transports (monotone schemes, detached inputs) run the value-only fused kernels"""
    import test_fv_dynamics  # noqa: F401  (makes sure the shared helpers are importable in isolation)
    test_step_api.test_step_api_reference_defaults_emu()


@pytest.mark.gpu
@pytest.mark.parametrize("params", [dict(hord=2), dict(hord=333), dict(hord=2, use_mf=1), dict(hord=1, n_sponge=1)])
def test_fused_vs_chain_gpu(monkeypatch, params):
    _fused_vs_chain(False, monkeypatch, 40, 2, params)


@pytest.mark.gpu
def test_fused_vs_chain_trajectory_schemes_gpu(monkeypatch):
    _fused_vs_chain(False, monkeypatch, 40, 2, dict(hord=9), tl=False)


@pytest.mark.gpu
def test_fused_step_gpu(fused):
    print(test_step_api._run(False, nonhydro=True))


# ---- row-marching forward kernels (csrc/fused_tp_march.h, FV3LM_TP_MARCH=1: opt-in, measured slower than the tile kernels) ----
@pytest.fixture
def marching(monkeypatch):
    common._handles.clear()
    monkeypatch.setenv("FV3LM_FUSED_TP", "2")
    monkeypatch.setenv("FV3LM_TP_MARCH", "1")
    yield
    common._handles.clear()


@pytest.mark.parametrize("hord", [1, 2, 333])
def test_march_vs_oracle_emu(marching, hord):
    test_tp_core._run(True, hord)


def test_march_vs_chain_emu(monkeypatch):
    """NL and TL of the marching kernels against the stage chain (N = 40: two row chunks of 24, both cube edges)"""
    monkeypatch.setenv("FV3LM_TP_MARCH", "1")
    _fused_vs_chain(True, monkeypatch, 40, 2, dict(hord=2))
    _fused_vs_chain(True, monkeypatch, 40, 2, dict(hord=2, use_mf=1, n_sponge=1))


def test_march_step_emu(marching):
    print(test_step_api._run(True, nonhydro=True))


# ---- a2b_ord4 as one tile kernel per direction (FV3LM_FUSED_A2B=1, csrc/a2b.cu: KernA2b / KernA2bRev) ----------------------
@pytest.fixture
def fused_all(monkeypatch):
    common._handles.clear()
    monkeypatch.setenv("FV3LM_FUSED_TP", "2")
    monkeypatch.setenv("FV3LM_FUSED_A2B", "1")
    monkeypatch.setenv("FV3LM_FUSED_CHAIN", "1")
    yield
    common._handles.clear()


@pytest.mark.parametrize("N", [12, 40])
def test_fused_a2b_vs_oracle_emu(fused_all, N):
    """NL, TL (= the operator itself), AD (= its transpose) against the oracle + dot-product test; N = 40 spans two tiles in x"""
    test_d_sw._run_a2b(True, N)


# ---- a2b_ord4 compiled into a constant stencil + sparse edge rows (FV3LM_FUSED_A2B=2, the default since round 2; csrc/a2b.cu a2bc) ----
@pytest.fixture
def a2b_compiled(monkeypatch):
    common._handles.clear()
    monkeypatch.setenv("FV3LM_FUSED_A2B", "2")
    yield
    common._handles.clear()


@pytest.mark.parametrize("N", [12, 40])
def test_compiled_a2b_vs_oracle_emu(a2b_compiled, N):
    """NL, TL, AD against the oracle + dot-product test; N = 40 has a wide regular rectangle, N = 12 is mostly edge rows"""
    test_d_sw._run_a2b(True, N)


def test_compiled_a2b_layout_2x2_emu(a2b_compiled):
    """sub-domains with zero, one or two cube edges: the regular rectangles and the sparse rows are built per resident sub-domain"""
    print(test_decomp._layout("a2b_ord4", True, (2, 2)))


def test_compiled_a2b_vs_chain_emu(monkeypatch):
    """the compiled operator against the stage chain it was probed from (NL and AD of the module, N = 40)"""
    res = {}
    for flag in ("0", "2"):
        common._handles.clear()
        monkeypatch.setenv("FV3LM_FUSED_A2B", flag)
        res[flag] = test_d_sw._a2b_nl_ad(True, 40)
    common._handles.clear()
    for a, b in zip(res["0"], res["2"]):
        assert np.abs(a).max() > 0 and relerr(b, a) <= 1e-13, relerr(b, a)


@pytest.mark.gpu
def test_compiled_a2b_vs_oracle_gpu(a2b_compiled):
    test_d_sw._run_a2b(False, 40)


@pytest.mark.parametrize("case", ["a2b_ord4", "fv_tp_2d", "step_nonhydro"])
def test_fused_layout_2x2_emu(fused_all, case):
    """tile kernels on layout(2, 2) sub-domains (cube-edge cases keyed on the tile-global index, 24 sub-domains in one launch)"""
    print(test_decomp._layout(case, True, (2, 2)))


def test_fused_all_step_emu(fused_all):
    print(test_step_api._run(True, nonhydro=True))


def test_fused_all_step_hydro_emu(fused_all):
    print(test_step_api._run(True, nonhydro=False))


def test_fused_multirank_gloo(fused_all):
    """two ranks over gloo (halo exchange + adjoint halo accumulation around the tile kernels): sharded = single-rank results"""
    import test_multirank
    test_multirank._check(2, True)


@pytest.mark.gpu
def test_fused_a2b_vs_oracle_gpu(fused_all):
    test_d_sw._run_a2b(False, 40)


@pytest.mark.gpu
def test_fused_all_step_gpu(fused_all):
    print(test_step_api._run(False, nonhydro=True))


# ---- automatic tile fusion of patch-free stage chains (FV3LM_FUSED_CHAIN=1, csrc/fused_chain.h): the tail of c_sw
# (cke, cvort, cwind) and the head of d_sw (dwind1, dwind2, dcourant, ra) in forward sweeps; `fused_all` switches it on as well ----
@pytest.mark.parametrize("case", ["c_sw", "d_sw", "dyn_core_nh_two_sided"])
def test_fused_chain_modules_emu(fused_all, case):
    """module level against the oracle (NL, TL, AD through the stage ops, dot product), whole tiles"""
    print(test_decomp.CASES[case](True))


@pytest.mark.parametrize("case", ["c_sw", "d_sw"])
def test_fused_chain_layout_2x2_emu(fused_all, case):
    print(test_decomp._layout(case, True, (2, 2)))


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["c_sw", "d_sw"])
def test_fused_chain_modules_gpu(fused_all, case):
    print(test_decomp.CASES[case](False))
