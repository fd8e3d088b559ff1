"""checks.dot_product_partial (what bench.py reports as `dot_product` at the bench workload) on a small case."""
import numpy as np
import pytest
import checks
import test_step_api as tsa


def _run(emu, nonhydro):
    h, f, comp, rng, cfg, ak, bk = tsa.make(emu, 12, 4, nonhydro=nonhydro)
    dx = {k: rng.standard_normal(comp[k].shape) * (np.abs(comp[k]).mean() * 1e-3) for k in tsa.ACT}
    keep = {k: v.copy() for k, v in dx.items()}
    lhs, rhs = checks.dot_product_partial(h, 0, dx, tsa.ACT)
    assert all(np.array_equal(dx[k], keep[k]) for k in dx)          # the caller's increment is not modified
    e = checks.rel_err(lhs, rhs)
    assert e < 1e-10, (lhs, rhs, e)
    return e


@pytest.mark.parametrize("nonhydro", [False, True])
def test_dot_product_check_emu(nonhydro):
    print(_run(True, nonhydro))


@pytest.mark.gpu
def test_dot_product_check_gpu():
    print(_run(False, True))
