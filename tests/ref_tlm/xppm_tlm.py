"""XPPM_TLM, model_tlmadm/tp_core_tlm.F90:2328-2489 (iord < 8 or 333 branch: the only one the tangent model has), transliterated."""
from . import F

p1, p2 = 7. / 12., -1. / 12.                 # tp_core_tlm.F90:40-41
c1, c2, c3 = -2. / 14., 11. / 14., 5. / 14.  # :53-55


def xppm_tlm(q, q_tl, c, c_tl, iord, is_, ie, isd, ied, jfirst, jlast, jsd, jed, npx, npy, dxa, nested=False, grid_type=0):
    """q, q_tl: F((isd,ied),(jfirst,jlast)); c, c_tl: F((is,ie+1),(jfirst,jlast)); dxa: F((isd,ied),(jsd,jed)).  Returns flux, flux_tl."""
    flux = F((is_, ie + 1), (jfirst, jlast)); flux_tl = F((is_, ie + 1), (jfirst, jlast))
    q1 = F((isd, ied)); q1_tl = F((isd, ied))
    al = F((is_ - 1, ie + 2)); al_tl = F((is_ - 1, ie + 2))
    if (not nested) and grid_type < 3:
        is1 = is_ - 1 if 3 < is_ - 1 else 3
        ie3 = ie + 2 if npx - 2 > ie + 2 else npx - 2
        ie1 = ie + 1 if npx - 3 > ie + 1 else npx - 3   # noqa: F841  (unused by the linear orders, kept from the source)
    else:
        is1 = is_ - 1; ie3 = ie + 2; ie1 = ie + 1          # noqa: F841
    al_tl.fill(0.0); q1_tl.fill(0.0)
    for j in range(jfirst, jlast + 1):
        for i in range(isd, ied + 1):
            q1_tl[i] = q_tl[i, j]
            q1[i] = q[i, j]
        if iord < 8 or iord == 333:
            for i in range(is1, ie3 + 1):
                al_tl[i] = p1 * (q1_tl[i - 1] + q1_tl[i]) + p2 * (q1_tl[i - 2] + q1_tl[i + 1])
                al[i] = p1 * (q1[i - 1] + q1[i]) + p2 * (q1[i - 2] + q1[i + 1])
            if (not nested) and grid_type < 3:
                if is_ == 1:
                    al_tl[0] = c1 * q1_tl[-2] + c2 * q1_tl[-1] + c3 * q1_tl[0]
                    al[0] = c1 * q1[-2] + c2 * q1[-1] + c3 * q1[0]
                    al_tl[1] = 0.5 * (((2. * dxa[0, j] + dxa[-1, j]) * q1_tl[0] - dxa[0, j] * q1_tl[-1]) / (dxa[-1, j] + dxa[0, j])
                                      + ((2. * dxa[1, j] + dxa[2, j]) * q1_tl[1] - dxa[1, j] * q1_tl[2]) / (dxa[1, j] + dxa[2, j]))
                    al[1] = 0.5 * (((2. * dxa[0, j] + dxa[-1, j]) * q1[0] - dxa[0, j] * q1[-1]) / (dxa[-1, j] + dxa[0, j])
                                   + ((2. * dxa[1, j] + dxa[2, j]) * q1[1] - dxa[1, j] * q1[2]) / (dxa[1, j] + dxa[2, j]))
                    al_tl[2] = c3 * q1_tl[1] + c2 * q1_tl[2] + c1 * q1_tl[3]
                    al[2] = c3 * q1[1] + c2 * q1[2] + c1 * q1[3]
                if ie + 1 == npx:
                    al_tl[npx - 1] = c1 * q1_tl[npx - 3] + c2 * q1_tl[npx - 2] + c3 * q1_tl[npx - 1]
                    al[npx - 1] = c1 * q1[npx - 3] + c2 * q1[npx - 2] + c3 * q1[npx - 1]
                    al_tl[npx] = 0.5 * (((2. * dxa[npx - 1, j] + dxa[npx - 2, j]) * q1_tl[npx - 1] - dxa[npx - 1, j] * q1_tl[npx - 2])
                                        / (dxa[npx - 2, j] + dxa[npx - 1, j])
                                        + ((2. * dxa[npx, j] + dxa[npx + 1, j]) * q1_tl[npx] - dxa[npx, j] * q1_tl[npx + 1])
                                        / (dxa[npx, j] + dxa[npx + 1, j]))
                    al[npx] = 0.5 * (((2. * dxa[npx - 1, j] + dxa[npx - 2, j]) * q1[npx - 1] - dxa[npx - 1, j] * q1[npx - 2])
                                     / (dxa[npx - 2, j] + dxa[npx - 1, j])
                                     + ((2. * dxa[npx, j] + dxa[npx + 1, j]) * q1[npx] - dxa[npx, j] * q1[npx + 1])
                                     / (dxa[npx, j] + dxa[npx + 1, j]))
                    al_tl[npx + 1] = c3 * q1_tl[npx] + c2 * q1_tl[npx + 1] + c1 * q1_tl[npx + 2]
                    al[npx + 1] = c3 * q1[npx] + c2 * q1[npx + 1] + c1 * q1[npx + 2]
            if iord == 1:
                for i in range(is_, ie + 2):
                    if c[i, j] > 0.:
                        flux_tl[i, j] = q1_tl[i - 1]
                        flux[i, j] = q1[i - 1]
                    else:
                        flux_tl[i, j] = q1_tl[i]
                        flux[i, j] = q1[i]
            elif iord == 2:
                for i in range(is_, ie + 2):
                    xt_tl = c_tl[i, j]
                    xt = c[i, j]
                    if xt > 0.:
                        qtmp_tl = q1_tl[i - 1]
                        qtmp = q1[i - 1]
                        flux_tl[i, j] = (qtmp_tl + (1. - xt) * (al_tl[i] - qtmp_tl - xt_tl * (al[i - 1] + al[i] - (qtmp + qtmp))
                                                                - xt * (al_tl[i - 1] + al_tl[i] - 2 * qtmp_tl))
                                         - xt_tl * (al[i] - qtmp - xt * (al[i - 1] + al[i] - (qtmp + qtmp))))
                        flux[i, j] = qtmp + (1. - xt) * (al[i] - qtmp - xt * (al[i - 1] + al[i] - (qtmp + qtmp)))
                    else:
                        qtmp_tl = q1_tl[i]
                        qtmp = q1[i]
                        flux_tl[i, j] = (qtmp_tl + xt_tl * (al[i] - qtmp + xt * (al[i] + al[i + 1] - (qtmp + qtmp)))
                                         + (1. + xt) * (al_tl[i] - qtmp_tl + xt_tl * (al[i] + al[i + 1] - (qtmp + qtmp))
                                                        + xt * (al_tl[i] + al_tl[i + 1] - 2 * qtmp_tl)))
                        flux[i, j] = qtmp + (1. + xt) * (al[i] - qtmp + xt * (al[i] + al[i + 1] - (qtmp + qtmp)))
            elif iord == 333:
                for i in range(is_, ie + 2):
                    xt_tl = c_tl[i, j]
                    xt = c[i, j]
                    if xt > 0.:
                        flux_tl[i, j] = ((2.0 * q1_tl[i] + 5.0 * q1_tl[i - 1] - q1_tl[i - 2]) / 6.0
                                         - 0.5 * (xt_tl * (q1[i] - q1[i - 1]) + xt * (q1_tl[i] - q1_tl[i - 1]))
                                         + (xt_tl * xt + xt * xt_tl) * (q1[i] - 2.0 * q1[i - 1] + q1[i - 2]) / 6.0
                                         + xt ** 2 * (q1_tl[i] - 2.0 * q1_tl[i - 1] + q1_tl[i - 2]) / 6.0)
                        flux[i, j] = ((2.0 * q1[i] + 5.0 * q1[i - 1] - q1[i - 2]) / 6.0 - 0.5 * xt * (q1[i] - q1[i - 1])
                                      + xt * xt / 6.0 * (q1[i] - 2.0 * q1[i - 1] + q1[i - 2]))
                    else:
                        flux_tl[i, j] = ((2.0 * q1_tl[i - 1] + 5.0 * q1_tl[i] - q1_tl[i + 1]) / 6.0
                                         - 0.5 * (xt_tl * (q1[i] - q1[i - 1]) + xt * (q1_tl[i] - q1_tl[i - 1]))
                                         + (xt_tl * xt + xt * xt_tl) * (q1[i + 1] - 2.0 * q1[i] + q1[i - 1]) / 6.0
                                         + xt ** 2 * (q1_tl[i + 1] - 2.0 * q1_tl[i] + q1_tl[i - 1]) / 6.0)
                        flux[i, j] = ((2.0 * q1[i - 1] + 5.0 * q1[i] - q1[i + 1]) / 6.0 - 0.5 * xt * (q1[i] - q1[i - 1])
                                      + xt * xt / 6.0 * (q1[i + 1] - 2.0 * q1[i] + q1[i - 1]))
    return flux, flux_tl
