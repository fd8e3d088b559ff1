"""EDGE_PROFILE_TLM (model_tlmadm/nh_utils_tlm.F90:3319-3472), transliterated: both the uniform-grid and the vertically varying branch,
and the top / bottom sign limiter.  Arrays are F((i1,i2),(j1,j2),(1,km[+1])); one call handles row j as in the source."""
from . import F


def edge_profile_tlm(q1, q1_tl, q2, q2_tl, q1e, q1e_tl, q2e, q2e_tl, i1, i2, j1, j2, j, km, dp0, uniform_grid, limiter):
    """dp0: F((1, km)).  q1e, q2e (and _tl) are filled in row j."""
    qe1 = F((i1, i2), (1, km + 1)); qe2 = F((i1, i2), (1, km + 1)); gam = F((i1, i2), (1, km + 1))
    qe1_tl = F((i1, i2), (1, km + 1)); qe2_tl = F((i1, i2), (1, km + 1))
    gak = F((1, km))
    if uniform_grid:
        r2o3 = 2. / 3.
        r4o3 = 4. / 3.
        for i in range(i1, i2 + 1):
            qe1_tl[i, 1] = r4o3 * q1_tl[i, j, 1] + r2o3 * q1_tl[i, j, 2]
            qe1[i, 1] = r4o3 * q1[i, j, 1] + r2o3 * q1[i, j, 2]
            qe2_tl[i, 1] = r4o3 * q2_tl[i, j, 1] + r2o3 * q2_tl[i, j, 2]
            qe2[i, 1] = r4o3 * q2[i, j, 1] + r2o3 * q2[i, j, 2]
        gak[1] = 7. / 3.
        for k in range(2, km + 1):
            gak[k] = 1. / (4. - gak[k - 1])
            for i in range(i1, i2 + 1):
                qe1_tl[i, k] = gak[k] * (3. * (q1_tl[i, j, k - 1] + q1_tl[i, j, k]) - qe1_tl[i, k - 1])
                qe1[i, k] = (3. * (q1[i, j, k - 1] + q1[i, j, k]) - qe1[i, k - 1]) * gak[k]
                qe2_tl[i, k] = gak[k] * (3. * (q2_tl[i, j, k - 1] + q2_tl[i, j, k]) - qe2_tl[i, k - 1])
                qe2[i, k] = (3. * (q2[i, j, k - 1] + q2[i, j, k]) - qe2[i, k - 1]) * gak[k]
        bet = 1. / (1.5 - 3.5 * gak[km])
        for i in range(i1, i2 + 1):
            qe1_tl[i, km + 1] = bet * (4. * q1_tl[i, j, km] + q1_tl[i, j, km - 1] - 3.5 * qe1_tl[i, km])
            qe1[i, km + 1] = (4. * q1[i, j, km] + q1[i, j, km - 1] - 3.5 * qe1[i, km]) * bet
            qe2_tl[i, km + 1] = bet * (4. * q2_tl[i, j, km] + q2_tl[i, j, km - 1] - 3.5 * qe2_tl[i, km])
            qe2[i, km + 1] = (4. * q2[i, j, km] + q2[i, j, km - 1] - 3.5 * qe2[i, km]) * bet
        for k in range(km, 0, -1):
            for i in range(i1, i2 + 1):
                qe1_tl[i, k] = qe1_tl[i, k] - gak[k] * qe1_tl[i, k + 1]
                qe1[i, k] = qe1[i, k] - gak[k] * qe1[i, k + 1]
                qe2_tl[i, k] = qe2_tl[i, k] - gak[k] * qe2_tl[i, k + 1]
                qe2[i, k] = qe2[i, k] - gak[k] * qe2[i, k + 1]
    else:
        g0 = dp0[2] / dp0[1]
        xt1 = 2. * g0 * (g0 + 1.)
        bet = g0 * (g0 + 0.5)
        for i in range(i1, i2 + 1):
            qe1_tl[i, 1] = (xt1 * q1_tl[i, j, 1] + q1_tl[i, j, 2]) / bet
            qe1[i, 1] = (xt1 * q1[i, j, 1] + q1[i, j, 2]) / bet
            qe2_tl[i, 1] = (xt1 * q2_tl[i, j, 1] + q2_tl[i, j, 2]) / bet
            qe2[i, 1] = (xt1 * q2[i, j, 1] + q2[i, j, 2]) / bet
            gam[i, 1] = (1. + g0 * (g0 + 1.5)) / bet
        for k in range(2, km + 1):
            gk = dp0[k - 1] / dp0[k]
            for i in range(i1, i2 + 1):
                bet = 2. + 2. * gk - gam[i, k - 1]
                qe1_tl[i, k] = (3. * (q1_tl[i, j, k - 1] + gk * q1_tl[i, j, k]) - qe1_tl[i, k - 1]) / bet
                qe1[i, k] = (3. * (q1[i, j, k - 1] + gk * q1[i, j, k]) - qe1[i, k - 1]) / bet
                qe2_tl[i, k] = (3. * (q2_tl[i, j, k - 1] + gk * q2_tl[i, j, k]) - qe2_tl[i, k - 1]) / bet
                qe2[i, k] = (3. * (q2[i, j, k - 1] + gk * q2[i, j, k]) - qe2[i, k - 1]) / bet
                gam[i, k] = gk / bet
        a_bot = 1. + gk * (gk + 1.5)
        xt1 = 2. * gk * (gk + 1.)
        for i in range(i1, i2 + 1):
            xt2 = gk * (gk + 0.5) - a_bot * gam[i, km]
            qe1_tl[i, km + 1] = (xt1 * q1_tl[i, j, km] + q1_tl[i, j, km - 1] - a_bot * qe1_tl[i, km]) / xt2
            qe1[i, km + 1] = (xt1 * q1[i, j, km] + q1[i, j, km - 1] - a_bot * qe1[i, km]) / xt2
            qe2_tl[i, km + 1] = (xt1 * q2_tl[i, j, km] + q2_tl[i, j, km - 1] - a_bot * qe2_tl[i, km]) / xt2
            qe2[i, km + 1] = (xt1 * q2[i, j, km] + q2[i, j, km - 1] - a_bot * qe2[i, km]) / xt2
        for k in range(km, 0, -1):
            for i in range(i1, i2 + 1):
                qe1_tl[i, k] = qe1_tl[i, k] - gam[i, k] * qe1_tl[i, k + 1]
                qe1[i, k] = qe1[i, k] - gam[i, k] * qe1[i, k + 1]
                qe2_tl[i, k] = qe2_tl[i, k] - gam[i, k] * qe2_tl[i, k + 1]
                qe2[i, k] = qe2[i, k] - gam[i, k] * qe2[i, k + 1]
    if limiter != 0:
        for i in range(i1, i2 + 1):
            if q1[i, j, 1] * qe1[i, 1] < 0.:
                qe1_tl[i, 1] = 0.0
                qe1[i, 1] = 0.
            if q2[i, j, 1] * qe2[i, 1] < 0.:
                qe2_tl[i, 1] = 0.0
                qe2[i, 1] = 0.
            if q1[i, j, km] * qe1[i, km + 1] < 0.:
                qe1_tl[i, km + 1] = 0.0
                qe1[i, km + 1] = 0.
            if q2[i, j, km] * qe2[i, km + 1] < 0.:
                qe2_tl[i, km + 1] = 0.0
                qe2[i, km + 1] = 0.
    for k in range(1, km + 2):
        for i in range(i1, i2 + 1):
            q1e_tl[i, j, k] = qe1_tl[i, k]
            q1e[i, j, k] = qe1[i, k]
            q2e_tl[i, j, k] = qe2_tl[i, k]
            q2e[i, j, k] = qe2[i, k]
