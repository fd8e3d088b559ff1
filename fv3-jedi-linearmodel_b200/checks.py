"""Acceptance checks through the public step API, usable at any size (bench.py reports the dot-product error of the bench
workload with it; BASELINE.json's metric names that error next to the throughput)."""
import numpy as np


def dot_product_partial(h, slot, dx, fields, seed=20261019):
    """<M dx, y> and <dx, M^T y> for the trajectory in `slot`, through fv3lm_step_tl / fv3lm_step_ad with host arrays.
    h: fv3lm.FV3LM; dx: dict of this rank's increment arrays.  Returns this rank's partial sums (lhs, rhs): add them over the
    ranks before comparing.  y is seeded noise scaled per field so that every field of M dx contributes O(1) to the sum."""
    a = {k: np.ascontiguousarray(dx[k], dtype=np.float64).copy() for k in fields}
    h.step_tl(slot, a)
    rng = np.random.default_rng(seed)
    y = {}
    for k in fields:
        s = float(np.abs(a[k]).mean())
        y[k] = rng.standard_normal(a[k].shape) / (s if s > 0.0 else 1.0)
    lhs = float(sum(np.vdot(a[k], y[k]) for k in fields))
    b = {k: y[k].copy() for k in fields}
    h.step_ad(slot, b)
    rhs = float(sum(np.vdot(dx[k], b[k]) for k in fields))
    return lhs, rhs


def rel_err(lhs, rhs):
    return abs(lhs - rhs) / max(abs(lhs), abs(rhs), 1e-300)
