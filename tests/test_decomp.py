"""Domain decomposition: every kernel family run on whole tiles and on tiles split into
layout(lx, ly) sub-domains (24 sub-domains batched on one rank) must agree for NL, TL and AD.  The
inter-rank transport itself is covered by tests/test_multirank.py (world_size 2, gloo)."""
import pytest
import common
import test_tp_core, test_c_sw, test_d_sw, test_dyn_core, test_fv_dynamics, test_nh, test_tracer_2d, test_c2l

CASES = {
    "fv_tp_2d": lambda emu: test_tp_core._run(emu, 2),
    "c_sw": lambda emu: test_c_sw._run(emu, False),
    "d_sw": lambda emu: test_d_sw._run_dsw(emu, False, False),
    "a2b_ord4": lambda emu: test_d_sw._run_a2b(emu),
    "dyn_core": lambda emu: test_dyn_core._run(emu, 2),
    "remap": lambda emu: test_fv_dynamics._run_remap(emu, True),
    "update_dz_c": lambda emu: test_nh._run_dzc(emu),
    "update_dz_d": lambda emu: test_nh._run_dzd(emu),
    "dyn_core_nh": lambda emu: test_nh._run_dyn_nh(emu, 2),
    "step_hydro": lambda emu: test_fv_dynamics._run_step(emu, 1, 2),
    "step_nonhydro": lambda emu: test_fv_dynamics._run_step(emu, 1, 2, nonhydro=True),
    "step_nonhydro_c24": lambda emu: test_fv_dynamics._run_step(emu, 1, 1, K=3, nonhydro=True, N=24),
    # two-sided mode (detached views, spliced chains, monotone trajectory schemes), the heat source path with del2_cubed
    "dyn_core_nh_two_sided": lambda emu: test_nh._run_dyn_nh(emu, 2, extra=test_dyn_core.TWO_SIDED),
    "dyn_core_two_sided_monotone": lambda emu: test_dyn_core._run(emu, 2, K=4, extra=test_dyn_core.TWO_SIDED_MONO),
    "dyn_core_heat": lambda emu: test_dyn_core._run(emu, 2, K=5, extra=dict(d_con=1.0)),
    "del2_cubed": lambda emu: test_dyn_core._run_del2_cubed(emu, 3),
    # q_split = 0: level maxima of the Courant numbers over 24 sub-domains, masked sub-steps with a halo update of q in between
    "c2l_ord4": lambda emu: test_c2l._run(emu),
    "tracer_2d_sub_steps": lambda emu: test_tracer_2d._run(emu, K=6, expect_nsplt=None),
}


def _layout(case, emu, layout):
    common.LAYOUT = layout
    try:
        return CASES[case](emu)
    finally:
        common.LAYOUT = None


@pytest.mark.parametrize("case", list(CASES))
def test_layout_2x2_emu(case):
    print(_layout(case, True, (2, 2)))


@pytest.mark.parametrize("case", ["d_sw", "step_nonhydro"])
def test_layout_1x2_emu(case):
    print(_layout(case, True, (1, 2)))


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["dyn_core", "step_nonhydro"])
def test_layout_2x2_gpu(case):
    _layout(case, False, (2, 2))
