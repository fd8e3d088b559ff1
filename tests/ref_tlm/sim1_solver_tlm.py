"""SIM1_SOLVER_TLM, model_tlmadm/nh_utils_tlm.F90:2548-2762, transliterated statement by statement (the `DO i` loops are numpy
vector statements over i; every `DO k` loop and the order of the statements are the source's).  Arrays are [km(+1), ni], Fortran
level k at row k - 1."""
import numpy as np

r3 = 1. / 3.        # nh_utils_tlm.F90:33


def sim1_solver_tlm(dt, km, rgas, gama, kappa, dm2, dm2_tl, pm2, pm2_tl, pem, pem_tl, w2, w2_tl, dz2, dz2_tl, pt2, pt2_tl, ws, ws_tl, p_fac):
    """returns pe, pe_tl (km+1), and the updated w2, w2_tl, dz2, dz2_tl (km)"""
    K = lambda k: k - 1
    ni = dm2.shape[1]
    w2 = w2.copy(); w2_tl = w2_tl.copy(); dz2 = dz2.copy(); dz2_tl = dz2_tl.copy()
    z = lambda n: np.zeros((n, ni))
    aa, bb, dd, w1, g_rat, gam = z(km), z(km), z(km), z(km), z(km), z(km)
    aa_tl, bb_tl, dd_tl, w1_tl, g_rat_tl, gam_tl = z(km), z(km), z(km), z(km), z(km), z(km)
    pp, pp_tl, pe, pe_tl = z(km + 1), z(km + 1), z(km + 1), z(km + 1)
    t1g = gama * 2. * dt * dt
    rdt = 1. / dt
    capa1 = kappa - 1.
    for k in range(1, km + 1):
        w1_tl[K(k)] = w2_tl[K(k)]
        w1[K(k)] = w2[K(k)]
        arg1_tl = -(rgas * ((dm2_tl[K(k)] * dz2[K(k)] - dm2[K(k)] * dz2_tl[K(k)]) * pt2[K(k)] / dz2[K(k)] ** 2 + dm2[K(k)] * pt2_tl[K(k)] / dz2[K(k)]))
        arg1 = -(dm2[K(k)] / dz2[K(k)] * rgas * pt2[K(k)])
        arg2_tl = gama * arg1_tl / arg1
        arg2 = gama * np.log(arg1)
        pe_tl[K(k)] = arg2_tl * np.exp(arg2) - pm2_tl[K(k)]
        pe[K(k)] = np.exp(arg2) - pm2[K(k)]
    for k in range(1, km):
        g_rat_tl[K(k)] = (dm2_tl[K(k)] * dm2[K(k + 1)] - dm2[K(k)] * dm2_tl[K(k + 1)]) / dm2[K(k + 1)] ** 2
        g_rat[K(k)] = dm2[K(k)] / dm2[K(k + 1)]
        bb_tl[K(k)] = 2. * g_rat_tl[K(k)]
        bb[K(k)] = 2. * (1. + g_rat[K(k)])
        dd_tl[K(k)] = 3. * (pe_tl[K(k)] + g_rat_tl[K(k)] * pe[K(k + 1)] + g_rat[K(k)] * pe_tl[K(k + 1)])
        dd[K(k)] = 3. * (pe[K(k)] + g_rat[K(k)] * pe[K(k + 1)])
    bet_tl = bb_tl[K(1)].copy()
    bet = bb[K(1)].copy()
    pp_tl[K(1)] = 0.0
    pp[K(1)] = 0.
    pp_tl[K(2)] = (dd_tl[K(1)] * bet - dd[K(1)] * bet_tl) / bet ** 2
    pp[K(2)] = dd[K(1)] / bet
    bb_tl[K(km)] = 0.0
    bb[K(km)] = 2.
    dd_tl[K(km)] = 3. * pe_tl[K(km)]
    dd[K(km)] = 3. * pe[K(km)]
    for k in range(2, km + 1):
        gam_tl[K(k)] = (g_rat_tl[K(k - 1)] * bet - g_rat[K(k - 1)] * bet_tl) / bet ** 2
        gam[K(k)] = g_rat[K(k - 1)] / bet
        bet_tl = bb_tl[K(k)] - gam_tl[K(k)]
        bet = bb[K(k)] - gam[K(k)]
        pp_tl[K(k + 1)] = ((dd_tl[K(k)] - pp_tl[K(k)]) * bet - (dd[K(k)] - pp[K(k)]) * bet_tl) / bet ** 2
        pp[K(k + 1)] = (dd[K(k)] - pp[K(k)]) / bet
    for k in range(km, 1, -1):
        pp_tl[K(k)] = pp_tl[K(k)] - gam_tl[K(k)] * pp[K(k + 1)] - gam[K(k)] * pp_tl[K(k + 1)]
        pp[K(k)] = pp[K(k)] - gam[K(k)] * pp[K(k + 1)]
    # Start the w-solver
    for k in range(2, km + 1):
        aa_tl[K(k)] = (t1g * (pem_tl[K(k)] + pp_tl[K(k)]) / (dz2[K(k - 1)] + dz2[K(k)])
                       - t1g * (dz2_tl[K(k - 1)] + dz2_tl[K(k)]) * (pem[K(k)] + pp[K(k)]) / (dz2[K(k - 1)] + dz2[K(k)]) ** 2)
        aa[K(k)] = t1g / (dz2[K(k - 1)] + dz2[K(k)]) * (pem[K(k)] + pp[K(k)])
    bet_tl = dm2_tl[K(1)] - aa_tl[K(2)]
    bet = dm2[K(1)] - aa[K(2)]
    w2_tl[K(1)] = ((dm2_tl[K(1)] * w1[K(1)] + dm2[K(1)] * w1_tl[K(1)] + dt * pp_tl[K(2)]) * bet - (dm2[K(1)] * w1[K(1)] + dt * pp[K(2)]) * bet_tl) / bet ** 2
    w2[K(1)] = (dm2[K(1)] * w1[K(1)] + dt * pp[K(2)]) / bet
    for k in range(2, km):
        gam_tl[K(k)] = (aa_tl[K(k)] * bet - aa[K(k)] * bet_tl) / bet ** 2
        gam[K(k)] = aa[K(k)] / bet
        bet_tl = dm2_tl[K(k)] - aa_tl[K(k)] - aa_tl[K(k + 1)] - aa_tl[K(k)] * gam[K(k)] - aa[K(k)] * gam_tl[K(k)]
        bet = dm2[K(k)] - (aa[K(k)] + aa[K(k + 1)] + aa[K(k)] * gam[K(k)])
        w2_tl[K(k)] = ((dm2_tl[K(k)] * w1[K(k)] + dm2[K(k)] * w1_tl[K(k)] + dt * (pp_tl[K(k + 1)] - pp_tl[K(k)]) - aa_tl[K(k)] * w2[K(k - 1)]
                        - aa[K(k)] * w2_tl[K(k - 1)]) * bet
                       - (dm2[K(k)] * w1[K(k)] + dt * (pp[K(k + 1)] - pp[K(k)]) - aa[K(k)] * w2[K(k - 1)]) * bet_tl) / bet ** 2
        w2[K(k)] = (dm2[K(k)] * w1[K(k)] + dt * (pp[K(k + 1)] - pp[K(k)]) - aa[K(k)] * w2[K(k - 1)]) / bet
    p1_tl = t1g * (pem_tl[K(km + 1)] + pp_tl[K(km + 1)]) / dz2[K(km)] - t1g * dz2_tl[K(km)] * (pem[K(km + 1)] + pp[K(km + 1)]) / dz2[K(km)] ** 2
    p1 = t1g / dz2[K(km)] * (pem[K(km + 1)] + pp[K(km + 1)])
    gam_tl[K(km)] = (aa_tl[K(km)] * bet - aa[K(km)] * bet_tl) / bet ** 2
    gam[K(km)] = aa[K(km)] / bet
    bet_tl = dm2_tl[K(km)] - aa_tl[K(km)] - p1_tl - aa_tl[K(km)] * gam[K(km)] - aa[K(km)] * gam_tl[K(km)]
    bet = dm2[K(km)] - (aa[K(km)] + p1 + aa[K(km)] * gam[K(km)])
    w2_tl[K(km)] = ((dm2_tl[K(km)] * w1[K(km)] + dm2[K(km)] * w1_tl[K(km)] + dt * (pp_tl[K(km + 1)] - pp_tl[K(km)]) - p1_tl * ws - p1 * ws_tl
                     - aa_tl[K(km)] * w2[K(km - 1)] - aa[K(km)] * w2_tl[K(km - 1)]) * bet
                    - (dm2[K(km)] * w1[K(km)] + dt * (pp[K(km + 1)] - pp[K(km)]) - p1 * ws - aa[K(km)] * w2[K(km - 1)]) * bet_tl) / bet ** 2
    w2[K(km)] = (dm2[K(km)] * w1[K(km)] + dt * (pp[K(km + 1)] - pp[K(km)]) - p1 * ws - aa[K(km)] * w2[K(km - 1)]) / bet
    for k in range(km - 1, 0, -1):
        w2_tl[K(k)] = w2_tl[K(k)] - gam_tl[K(k + 1)] * w2[K(k + 1)] - gam[K(k + 1)] * w2_tl[K(k + 1)]
        w2[K(k)] = w2[K(k)] - gam[K(k + 1)] * w2[K(k + 1)]
    pe_tl[K(1)] = 0.0
    pe[K(1)] = 0.
    for k in range(1, km + 1):
        pe_tl[K(k + 1)] = pe_tl[K(k)] + rdt * (dm2_tl[K(k)] * (w2[K(k)] - w1[K(k)]) + dm2[K(k)] * (w2_tl[K(k)] - w1_tl[K(k)]))
        pe[K(k + 1)] = pe[K(k)] + dm2[K(k)] * (w2[K(k)] - w1[K(k)]) * rdt
    p1_tl = r3 * (pe_tl[K(km)] + 2. * pe_tl[K(km + 1)])
    p1 = (pe[K(km)] + 2. * pe[K(km + 1)]) * r3
    first = p_fac * pm2[K(km)] < p1 + pm2[K(km)]
    max1_tl = np.where(first, p1_tl + pm2_tl[K(km)], p_fac * pm2_tl[K(km)])
    max1 = np.where(first, p1 + pm2[K(km)], p_fac * pm2[K(km)])
    arg1_tl = capa1 * max1_tl / max1
    arg1 = capa1 * np.log(max1)
    dz2_tl[K(km)] = -(rgas * ((dm2_tl[K(km)] * pt2[K(km)] + dm2[K(km)] * pt2_tl[K(km)]) * np.exp(arg1) + dm2[K(km)] * pt2[K(km)] * arg1_tl * np.exp(arg1)))
    dz2[K(km)] = -(dm2[K(km)] * rgas * pt2[K(km)] * np.exp(arg1))
    for k in range(km - 1, 0, -1):
        p1_tl = (r3 * (pe_tl[K(k)] + bb_tl[K(k)] * pe[K(k + 1)] + bb[K(k)] * pe_tl[K(k + 1)] + g_rat_tl[K(k)] * pe[K(k + 2)] + g_rat[K(k)] * pe_tl[K(k + 2)])
                 - g_rat_tl[K(k)] * p1 - g_rat[K(k)] * p1_tl)
        p1 = (pe[K(k)] + bb[K(k)] * pe[K(k + 1)] + g_rat[K(k)] * pe[K(k + 2)]) * r3 - g_rat[K(k)] * p1
        first = p_fac * pm2[K(k)] < p1 + pm2[K(k)]
        max2_tl = np.where(first, p1_tl + pm2_tl[K(k)], p_fac * pm2_tl[K(k)])
        max2 = np.where(first, p1 + pm2[K(k)], p_fac * pm2[K(k)])
        arg1_tl = capa1 * max2_tl / max2
        arg1 = capa1 * np.log(max2)
        dz2_tl[K(k)] = -(rgas * ((dm2_tl[K(k)] * pt2[K(k)] + dm2[K(k)] * pt2_tl[K(k)]) * np.exp(arg1) + dm2[K(k)] * pt2[K(k)] * arg1_tl * np.exp(arg1)))
        dz2[K(k)] = -(dm2[K(k)] * rgas * pt2[K(k)] * np.exp(arg1))
    return pe, pe_tl, w2, w2_tl, dz2, dz2_tl
