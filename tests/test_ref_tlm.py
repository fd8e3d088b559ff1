"""Reference-derived pin of the ORACLE: the oracle's tangent (torch.func.jvp of its restated primal) against
statement-by-statement numpy transliterations of the reference's own Tapenade tangent routines (tests/ref_tlm/).  CPU only."""
import numpy as np
import pytest
import torch
from oracle import tp_core as otp
from common import metrics, ograd, rnd, relerr, region
from ref_tlm import F
from ref_tlm.xppm_tlm import xppm_tlm

TOL = 1e-13    # same arithmetic up to the order of a few additions


def _fa(a2d, N):
    """[NY, NX] array (Fortran (i, j) at [j + 2, i + 2]) -> F((isd, ied + 1), (jsd, jed + 1)) view indexed (i, j)"""
    return F((-2, N + 4), (-2, N + 4), data=np.ascontiguousarray(a2d.T))


@pytest.mark.parametrize("iord", [1, 2, 333])
def test_xppm_tlm_pins_oracle(iord):
    """XPPM_TLM (model_tlmadm/tp_core_tlm.F90:2328-2489) on whole cube tiles: is = 1, ie = N = npx - 1, both cube edges present"""
    N, K = 12, 2
    rng = np.random.default_rng(5)
    g = ograd(N)
    q = rnd(rng, N, K, 1.0, 10.0); c = rnd(rng, N, K, 0.4)
    dq = rnd(rng, N, K, 0.1); dc = rnd(rng, N, K, 0.01)
    j0, j1 = -2, N + 3                                   # jfirst:jlast = jsd:jed (the inner sweep of fv_tp_2d)
    fn = lambda a, b: otp.xppm(a, b, iord, g, j0, j1)
    flux_o, dflux_o = torch.func.jvp(fn, (torch.from_numpy(q), torch.from_numpy(c)), (torch.from_numpy(dq), torch.from_numpy(dc)))
    dxa = metrics(N)["dxa"]
    worst = 0.0
    for t in range(6):
        for k in range(K):
            fq, fq_tl, fc, fc_tl = _fa(q[t, k], N), _fa(dq[t, k], N), _fa(c[t, k], N), _fa(dc[t, k], N)
            # q(isd:ied, jfirst:jlast), c(is:ie+1, ...): the wider F arrays hold them at the same Fortran indices
            flux, flux_tl = xppm_tlm(fq, fq_tl, fc, fc_tl, iord, 1, N, -2, N + 3, j0, j1, -2, N + 3, N + 1, N + 1, _fa(dxa[t], N))
            ref = flux.a.T          # [j - j0, i - 1]
            ref_tl = flux_tl.a.T
            o = region(flux_o[t, k].numpy(), 1, N + 1, j0, j1)
            o_tl = region(dflux_o[t, k].numpy(), 1, N + 1, j0, j1)
            worst = max(worst, relerr(o, ref), relerr(o_tl, ref_tl))
            assert np.abs(ref_tl).max() > 0
    assert worst <= TOL, worst


def test_sim1_solver_tlm_pins_oracle():
    """SIM1_SOLVER_TLM (model_tlmadm/nh_utils_tlm.F90:2548-2762) on 40 columns of 72 layers with atmosphere-like magnitudes"""
    from oracle import nh as onh
    from ref_tlm.sim1_solver_tlm import sim1_solver_tlm
    km, ni = 72, 40
    rng = np.random.default_rng(11)
    rgas, kappa = 287.05, 287.05 / 1004.6
    gama = 1. / (1. - kappa)
    dt, p_fac = 25.0, 0.05
    dp = np.linspace(2.0, 1600.0, km)[:, None] * (1.0 + 0.05 * rng.standard_normal((km, ni)))      # Pa
    dm2 = dp / 9.80665
    pem = np.concatenate([np.full((1, ni), 1.0), 1.0 + np.cumsum(dp, axis=0)], axis=0)
    pm2 = (pem[1:] - pem[:-1]) / np.log(pem[1:] / pem[:-1])
    pt2 = (300.0 + 5.0 * rng.standard_normal((km, ni)))
    dz2 = -(dm2 * rgas * pt2 / pm2 ** (1. - kappa) * pm2 ** (-kappa) * 1.0) * (pm2 ** (kappa) / pm2 ** kappa)   # ~ -dm R T / p (hydrostatic)
    dz2 = dz2 * (1.0 + 0.01 * rng.standard_normal((km, ni)))
    w2 = 0.5 * rng.standard_normal((km, ni))
    ws = 0.1 * rng.standard_normal(ni)
    pert = lambda a, s: s * np.abs(a).mean() * rng.standard_normal(a.shape)
    d = dict(dm2=pert(dm2, 1e-3), pm2=pert(pm2, 1e-3), pem=pert(pem, 1e-3), w2=pert(w2, 1e-1), dz2=pert(dz2, 1e-3), pt2=pert(pt2, 1e-3), ws=pert(ws, 1e-1))
    pe, pe_tl, w2n, w2n_tl, dzn, dzn_tl = sim1_solver_tlm(dt, km, rgas, gama, kappa, dm2, d["dm2"], pm2, d["pm2"], pem, d["pem"], w2, d["w2"],
                                                          dz2, d["dz2"], pt2, d["pt2"], ws, d["ws"], p_fac)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))

    def fn(dm2_, pm2_, pem_, w2_, dz2_, pt2_, ws_):
        w, p, z = onh.sim1_solver(dt, list(dm2_), list(pm2_), list(pem_), list(w2_), list(dz2_), list(pt2_), ws_, rgas, gama, kappa, p_fac)
        return torch.stack(w), torch.stack(p), torch.stack(z)
    prim = (T(dm2), T(pm2), T(pem), T(w2), T(dz2), T(pt2), T(ws))
    tang = (T(d["dm2"]), T(d["pm2"]), T(d["pem"]), T(d["w2"]), T(d["dz2"]), T(d["pt2"]), T(d["ws"]))
    (w_o, p_o, z_o), (dw_o, dp_o, dz_o) = torch.func.jvp(fn, prim, tang)
    errs = dict(w2=relerr(w_o.numpy(), w2n), pe=relerr(p_o.numpy(), pe), dz2=relerr(z_o.numpy(), dzn),
                w2_tl=relerr(dw_o.numpy(), w2n_tl), pe_tl=relerr(dp_o.numpy(), pe_tl), dz2_tl=relerr(dz_o.numpy(), dzn_tl))
    assert np.abs(w2n_tl).max() > 0 and np.abs(dzn_tl).max() > 0
    print('sim1 errs', errs)
    assert max(errs.values()) <= 1e-12, errs      # achieved 1e-13 (exp / log round-off through 72 levels)


def test_a2b_ord4_tlm_pins_oracle():
    """A2B_ORD4_TLM (model_tlmadm/a2b_edge_tlm.F90:546-1158, with EXTRAP_CORNER_TLM :1491) on whole cube tiles: all four corners, all four
    edges, the interior; the operator is linear, so the oracle's jvp must be the operator applied to the perturbation"""
    from oracle.a2b_edge import a2b_ord4
    from ref_tlm.a2b_ord4_tlm import a2b_ord4_tlm
    N, K = 12, 1
    rng = np.random.default_rng(23)
    g = ograd(N)
    M = metrics(N)
    q = rnd(rng, N, K, 1.0, 5.0); dq = rnd(rng, N, K, 0.3)
    fn = lambda a: a2b_ord4(a, g)
    out_o, dout_o = torch.func.jvp(fn, (torch.from_numpy(q),), (torch.from_numpy(dq),))
    npx = npy = N + 1
    worst = 0.0
    for t in range(6):
        P = lambda a, i, j: a[t, j + 2, i + 2]                # Fortran (i, j) of a [6, NY, NX(, 2)] metric
        qout, qout_tl = a2b_ord4_tlm(_fa(q[t, 0], N), _fa(dq[t, 0], N), lambda i, j: P(M["grid"], i, j), lambda i, j: P(M["agrid"], i, j),
                                     _fa(M["dxa"][t], N), _fa(M["dya"][t], N),
                                     lambda j: M["edge_w"][t, j + 2], lambda j: M["edge_e"][t, j + 2], lambda i: M["edge_s"][t, i + 2], lambda i: M["edge_n"][t, i + 2],
                                     npx, npy, 1, N, 1, N, 3)
        ref = qout.a.T; ref_tl = qout_tl.a.T                   # [j + 2, i + 2]
        o = region(out_o[t, 0].numpy(), 1, N + 1, 1, N + 1); o_tl = region(dout_o[t, 0].numpy(), 1, N + 1, 1, N + 1)
        worst = max(worst, relerr(o, region(ref, 1, N + 1, 1, N + 1)), relerr(o_tl, region(ref_tl, 1, N + 1, 1, N + 1)))
        assert np.abs(region(ref_tl, 1, N + 1, 1, N + 1)).min() > 0        # every corner point of the tile is defined
    print("a2b_ord4 worst", worst)
    assert worst <= TOL, worst
