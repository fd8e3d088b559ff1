"""FV_TP_2D_TLM (model_tlmadm/tp_core_tlm.F90:2123-2324) with the routines it calls: YPPM_TLM (:2496-2669), DELN_FLUX_TLM (:2673-2836) and
COPY_CORNERS_TLM (:2843-2924), transliterated statement by statement (whole cube tile: all four corners present, not nested, grid_type 0).
XPPM_TLM is in xppm_tlm.py."""
from . import F
from .xppm_tlm import xppm_tlm, p1, p2, c1, c2, c3

ng = 3      # fv_mp halo width used by copy_corners


class BD:
    """fv_grid_bounds_type of a whole tile"""

    def __init__(self, N):
        self.is_ = 1; self.ie = N; self.js = 1; self.je = N
        self.isd = 1 - ng; self.ied = N + ng; self.jsd = 1 - ng; self.jed = N + ng


def copy_corners_tlm(q, q_tl, npx, npy, dir_):
    """:2843-2924, sw / se / ne / nw corners all present"""
    if dir_ == 1:
        for j in range(1 - ng, 1):
            for i in range(1 - ng, 1):
                q_tl[i, j] = q_tl[j, 1 - i]
                q[i, j] = q[j, 1 - i]
        for j in range(1 - ng, 1):
            for i in range(npx, npx + ng):
                q_tl[i, j] = q_tl[npy - j, i - npx + 1]
                q[i, j] = q[npy - j, i - npx + 1]
        for j in range(npy, npy + ng):
            for i in range(npx, npx + ng):
                q_tl[i, j] = q_tl[j, 2 * npx - 1 - i]
                q[i, j] = q[j, 2 * npx - 1 - i]
        for j in range(npy, npy + ng):
            for i in range(1 - ng, 1):
                q_tl[i, j] = q_tl[npy - j, i - 1 + npx]
                q[i, j] = q[npy - j, i - 1 + npx]
    elif dir_ == 2:
        for j in range(1 - ng, 1):
            for i in range(1 - ng, 1):
                q_tl[i, j] = q_tl[1 - j, i]
                q[i, j] = q[1 - j, i]
        for j in range(1 - ng, 1):
            for i in range(npx, npx + ng):
                q_tl[i, j] = q_tl[npy + j - 1, npx - i]
                q[i, j] = q[npy + j - 1, npx - i]
        for j in range(npy, npy + ng):
            for i in range(npx, npx + ng):
                q_tl[i, j] = q_tl[2 * npy - 1 - j, i]
                q[i, j] = q[2 * npy - 1 - j, i]
        for j in range(npy, npy + ng):
            for i in range(1 - ng, 1):
                q_tl[i, j] = q_tl[j + 1 - npx, npy - i]
                q[i, j] = q[j + 1 - npx, npy - i]


def yppm_tlm(q, q_tl, c, c_tl, jord, ifirst, ilast, isd, ied, js, je, jsd, jed, npx, npy, dya):
    """:2496-2669.  q(ifirst:ilast, jsd:jed), c(isd:ied, js:je+1) -> flux(ifirst:ilast, js:je+1)"""
    flux = F((ifirst, ilast), (js, je + 1)); flux_tl = F((ifirst, ilast), (js, je + 1))
    al = F((ifirst, ilast), (js - 1, je + 2)); al_tl = F((ifirst, ilast), (js - 1, je + 2))
    js1 = js - 1 if 3 < js - 1 else 3
    je3 = je + 2 if npy - 2 > je + 2 else npy - 2
    assert jord < 8 or jord == 333
    for j in range(js1, je3 + 1):
        for i in range(ifirst, ilast + 1):
            al_tl[i, j] = p1 * (q_tl[i, j - 1] + q_tl[i, j]) + p2 * (q_tl[i, j - 2] + q_tl[i, j + 1])
            al[i, j] = p1 * (q[i, j - 1] + q[i, j]) + p2 * (q[i, j - 2] + q[i, j + 1])
    if js == 1:
        for i in range(ifirst, ilast + 1):
            al_tl[i, 0] = c1 * q_tl[i, -2] + c2 * q_tl[i, -1] + c3 * q_tl[i, 0]
            al[i, 0] = c1 * q[i, -2] + c2 * q[i, -1] + c3 * q[i, 0]
            al_tl[i, 1] = 0.5 * (((2. * dya[i, 0] + dya[i, -1]) * q_tl[i, 0] - dya[i, 0] * q_tl[i, -1]) / (dya[i, -1] + dya[i, 0])
                                 + ((2. * dya[i, 1] + dya[i, 2]) * q_tl[i, 1] - dya[i, 1] * q_tl[i, 2]) / (dya[i, 1] + dya[i, 2]))
            al[i, 1] = 0.5 * (((2. * dya[i, 0] + dya[i, -1]) * q[i, 0] - dya[i, 0] * q[i, -1]) / (dya[i, -1] + dya[i, 0])
                              + ((2. * dya[i, 1] + dya[i, 2]) * q[i, 1] - dya[i, 1] * q[i, 2]) / (dya[i, 1] + dya[i, 2]))
            al_tl[i, 2] = c3 * q_tl[i, 1] + c2 * q_tl[i, 2] + c1 * q_tl[i, 3]
            al[i, 2] = c3 * q[i, 1] + c2 * q[i, 2] + c1 * q[i, 3]
    if je + 1 == npy:
        for i in range(ifirst, ilast + 1):
            al_tl[i, npy - 1] = c1 * q_tl[i, npy - 3] + c2 * q_tl[i, npy - 2] + c3 * q_tl[i, npy - 1]
            al[i, npy - 1] = c1 * q[i, npy - 3] + c2 * q[i, npy - 2] + c3 * q[i, npy - 1]
            al_tl[i, npy] = 0.5 * (((2. * dya[i, npy - 1] + dya[i, npy - 2]) * q_tl[i, npy - 1] - dya[i, npy - 1] * q_tl[i, npy - 2])
                                   / (dya[i, npy - 2] + dya[i, npy - 1])
                                   + ((2. * dya[i, npy] + dya[i, npy + 1]) * q_tl[i, npy] - dya[i, npy] * q_tl[i, npy + 1])
                                   / (dya[i, npy] + dya[i, npy + 1]))
            al[i, npy] = 0.5 * (((2. * dya[i, npy - 1] + dya[i, npy - 2]) * q[i, npy - 1] - dya[i, npy - 1] * q[i, npy - 2])
                                / (dya[i, npy - 2] + dya[i, npy - 1])
                                + ((2. * dya[i, npy] + dya[i, npy + 1]) * q[i, npy] - dya[i, npy] * q[i, npy + 1])
                                / (dya[i, npy] + dya[i, npy + 1]))
            al_tl[i, npy + 1] = c3 * q_tl[i, npy] + c2 * q_tl[i, npy + 1] + c1 * q_tl[i, npy + 2]
            al[i, npy + 1] = c3 * q[i, npy] + c2 * q[i, npy + 1] + c1 * q[i, npy + 2]
    if jord == 1:
        for j in range(js, je + 2):
            for i in range(ifirst, ilast + 1):
                if c[i, j] > 0.:
                    flux_tl[i, j] = q_tl[i, j - 1]
                    flux[i, j] = q[i, j - 1]
                else:
                    flux_tl[i, j] = q_tl[i, j]
                    flux[i, j] = q[i, j]
    elif jord == 2:
        for j in range(js, je + 2):
            for i in range(ifirst, ilast + 1):
                xt_tl = c_tl[i, j]
                xt = c[i, j]
                if xt > 0.:
                    qtmp_tl = q_tl[i, j - 1]
                    qtmp = q[i, j - 1]
                    flux_tl[i, j] = (qtmp_tl + (1. - xt) * (al_tl[i, j] - qtmp_tl - xt_tl * (al[i, j - 1] + al[i, j] - (qtmp + qtmp))
                                                            - xt * (al_tl[i, j - 1] + al_tl[i, j] - 2 * qtmp_tl))
                                     - xt_tl * (al[i, j] - qtmp - xt * (al[i, j - 1] + al[i, j] - (qtmp + qtmp))))
                    flux[i, j] = qtmp + (1. - xt) * (al[i, j] - qtmp - xt * (al[i, j - 1] + al[i, j] - (qtmp + qtmp)))
                else:
                    qtmp_tl = q_tl[i, j]
                    qtmp = q[i, j]
                    flux_tl[i, j] = (qtmp_tl + xt_tl * (al[i, j] - qtmp + xt * (al[i, j] + al[i, j + 1] - (qtmp + qtmp)))
                                     + (1. + xt) * (al_tl[i, j] - qtmp_tl + xt_tl * (al[i, j] + al[i, j + 1] - (qtmp + qtmp))
                                                    + xt * (al_tl[i, j] + al_tl[i, j + 1] - 2 * qtmp_tl)))
                    flux[i, j] = qtmp + (1. + xt) * (al[i, j] - qtmp + xt * (al[i, j] + al[i, j + 1] - (qtmp + qtmp)))
    elif jord == 333:
        for j in range(js, je + 2):
            for i in range(ifirst, ilast + 1):
                xt_tl = c_tl[i, j]
                xt = c[i, j]
                if xt > 0.:
                    flux_tl[i, j] = ((2.0 * q_tl[i, j] + 5.0 * q_tl[i, j - 1] - q_tl[i, j - 2]) / 6.0
                                     - 0.5 * (xt_tl * (q[i, j] - q[i, j - 1]) + xt * (q_tl[i, j] - q_tl[i, j - 1]))
                                     + (xt_tl * xt + xt * xt_tl) * (q[i, j] - 2.0 * q[i, j - 1] + q[i, j - 2]) / 6.0
                                     + xt ** 2 * (q_tl[i, j] - 2.0 * q_tl[i, j - 1] + q_tl[i, j - 2]) / 6.0)
                    flux[i, j] = ((2.0 * q[i, j] + 5.0 * q[i, j - 1] - q[i, j - 2]) / 6.0 - 0.5 * xt * (q[i, j] - q[i, j - 1])
                                  + xt * xt / 6.0 * (q[i, j] - 2.0 * q[i, j - 1] + q[i, j - 2]))
                else:
                    flux_tl[i, j] = ((2.0 * q_tl[i, j - 1] + 5.0 * q_tl[i, j] - q_tl[i, j + 1]) / 6.0
                                     - 0.5 * (xt_tl * (q[i, j] - q[i, j - 1]) + xt * (q_tl[i, j] - q_tl[i, j - 1]))
                                     + (xt_tl * xt + xt * xt_tl) * (q[i, j + 1] - 2.0 * q[i, j] + q[i, j - 1]) / 6.0
                                     + xt ** 2 * (q_tl[i, j + 1] - 2.0 * q_tl[i, j] + q_tl[i, j - 1]) / 6.0)
                    flux[i, j] = ((2.0 * q[i, j - 1] + 5.0 * q[i, j] - q[i, j + 1]) / 6.0 - 0.5 * xt * (q[i, j] - q[i, j - 1])
                                  + xt * xt / 6.0 * (q[i, j + 1] - 2.0 * q[i, j] + q[i, j - 1]))
    return flux, flux_tl


def deln_flux_tlm(nord, is_, ie, js, je, npx, npy, damp, q, q_tl, fx, fx_tl, fy, fy_tl, gs, bd, mass=None, mass_tl=None):
    """:2673-2836.  fx, fy (and _tl) updated in place.  gs: dict of F arrays del6_u, del6_v, rarea"""
    fx2 = F((bd.isd, bd.ied + 1), (bd.jsd, bd.jed)); fx2_tl = F((bd.isd, bd.ied + 1), (bd.jsd, bd.jed))
    fy2 = F((bd.isd, bd.ied), (bd.jsd, bd.jed + 1)); fy2_tl = F((bd.isd, bd.ied), (bd.jsd, bd.jed + 1))
    d2 = F((bd.isd, bd.ied), (bd.jsd, bd.jed)); d2_tl = F((bd.isd, bd.ied), (bd.jsd, bd.jed))
    del6_u, del6_v, rarea = gs["del6_u"], gs["del6_v"], gs["rarea"]
    i1 = is_ - 1 - nord
    i2 = ie + 1 + nord
    j1 = js - 1 - nord
    j2 = je + 1 + nord
    if mass is None:
        for j in range(j1, j2 + 1):
            for i in range(i1, i2 + 1):
                d2_tl[i, j] = damp * q_tl[i, j]
                d2[i, j] = damp * q[i, j]
    else:
        for j in range(j1, j2 + 1):
            for i in range(i1, i2 + 1):
                d2_tl[i, j] = q_tl[i, j]
                d2[i, j] = q[i, j]
    if nord > 0:
        copy_corners_tlm(d2, d2_tl, npx, npy, 1)
    for j in range(js - nord, je + nord + 1):
        for i in range(is_ - nord, ie + nord + 2):
            fx2_tl[i, j] = del6_v[i, j] * (d2_tl[i - 1, j] - d2_tl[i, j])
            fx2[i, j] = del6_v[i, j] * (d2[i - 1, j] - d2[i, j])
    if nord > 0:
        copy_corners_tlm(d2, d2_tl, npx, npy, 2)
    for j in range(js - nord, je + nord + 2):
        for i in range(is_ - nord, ie + nord + 1):
            fy2_tl[i, j] = del6_u[i, j] * (d2_tl[i, j - 1] - d2_tl[i, j])
            fy2[i, j] = del6_u[i, j] * (d2[i, j - 1] - d2[i, j])
    if nord > 0:
        for n in range(1, nord + 1):
            nt = nord - n
            for j in range(js - nt - 1, je + nt + 2):
                for i in range(is_ - nt - 1, ie + nt + 2):
                    d2_tl[i, j] = rarea[i, j] * (fx2_tl[i, j] - fx2_tl[i + 1, j] + fy2_tl[i, j] - fy2_tl[i, j + 1])
                    d2[i, j] = (fx2[i, j] - fx2[i + 1, j] + fy2[i, j] - fy2[i, j + 1]) * rarea[i, j]
            copy_corners_tlm(d2, d2_tl, npx, npy, 1)
            for j in range(js - nt, je + nt + 1):
                for i in range(is_ - nt, ie + nt + 2):
                    fx2_tl[i, j] = del6_v[i, j] * (d2_tl[i, j] - d2_tl[i - 1, j])
                    fx2[i, j] = del6_v[i, j] * (d2[i, j] - d2[i - 1, j])
            copy_corners_tlm(d2, d2_tl, npx, npy, 2)
            for j in range(js - nt, je + nt + 2):
                for i in range(is_ - nt, ie + nt + 1):
                    fy2_tl[i, j] = del6_u[i, j] * (d2_tl[i, j] - d2_tl[i, j - 1])
                    fy2[i, j] = del6_u[i, j] * (d2[i, j] - d2[i, j - 1])
    if mass is not None:
        damp2 = 0.5 * damp
        for j in range(js, je + 1):
            for i in range(is_, ie + 2):
                fx_tl[i, j] = fx_tl[i, j] + damp2 * ((mass_tl[i - 1, j] + mass_tl[i, j]) * fx2[i, j] + (mass[i - 1, j] + mass[i, j]) * fx2_tl[i, j])
                fx[i, j] = fx[i, j] + damp2 * (mass[i - 1, j] + mass[i, j]) * fx2[i, j]
        for j in range(js, je + 2):
            for i in range(is_, ie + 1):
                fy_tl[i, j] = fy_tl[i, j] + damp2 * ((mass_tl[i, j - 1] + mass_tl[i, j]) * fy2[i, j] + (mass[i, j - 1] + mass[i, j]) * fy2_tl[i, j])
                fy[i, j] = fy[i, j] + damp2 * (mass[i, j - 1] + mass[i, j]) * fy2[i, j]
    else:
        for j in range(js, je + 1):
            for i in range(is_, ie + 2):
                fx_tl[i, j] = fx_tl[i, j] + fx2_tl[i, j]
                fx[i, j] = fx[i, j] + fx2[i, j]
        for j in range(js, je + 2):
            for i in range(is_, ie + 1):
                fy_tl[i, j] = fy_tl[i, j] + fy2_tl[i, j]
                fy[i, j] = fy[i, j] + fy2[i, j]


def _sub(a, i0, i1, j0, j1):
    """the array section a(i0:i1, j0:j1) passed as an actual argument: a new F with those bounds"""
    s = F((i0, i1), (j0, j1))
    for j in range(j0, j1 + 1):
        for i in range(i0, i1 + 1):
            s[i, j] = a[i, j]
    return s


def fv_tp_2d_tlm(q, q_tl, crx, crx_tl, cry, cry_tl, npx, npy, hord, xfx, xfx_tl, yfx, yfx_tl, gs, bd, ra_x, ra_x_tl, ra_y, ra_y_tl,
                 mfx=None, mfx_tl=None, mfy=None, mfy_tl=None, mass=None, mass_tl=None, nord=None, damp_c=None):
    """:2123-2324.  q, q_tl are modified in place (copy_corners) as in the source.  gs: dict of F arrays area, rarea, dxa, dya, del6_u,
    del6_v and the scalar da_min.  Returns fx, fx_tl (is:ie+1, js:je), fy, fy_tl (is:ie, js:je+1)."""
    is_, ie, js, je = bd.is_, bd.ie, bd.js, bd.je
    isd, ied, jsd, jed = bd.isd, bd.ied, bd.jsd, bd.jed
    area = gs["area"]
    ord_in = 8 if hord == 10 else hord
    ord_ou = hord
    q_i = F((isd, ied), (js, je)); q_i_tl = F((isd, ied), (js, je))
    q_j = F((is_, ie), (jsd, jed)); q_j_tl = F((is_, ie), (jsd, jed))
    fyy = F((isd, ied), (js, je + 1)); fyy_tl = F((isd, ied), (js, je + 1))
    fx1 = F((is_, ie + 1)); fx1_tl = F((is_, ie + 1))
    copy_corners_tlm(q, q_tl, npx, npy, 2)
    fy2, fy2_tl = yppm_tlm(q, q_tl, cry, cry_tl, ord_in, isd, ied, isd, ied, js, je, jsd, jed, npx, npy, gs["dya"])
    for j in range(js, je + 2):
        for i in range(isd, ied + 1):
            fyy_tl[i, j] = yfx_tl[i, j] * fy2[i, j] + yfx[i, j] * fy2_tl[i, j]
            fyy[i, j] = yfx[i, j] * fy2[i, j]
    for j in range(js, je + 1):
        for i in range(isd, ied + 1):
            q_i_tl[i, j] = (((area[i, j] * q_tl[i, j] + fyy_tl[i, j] - fyy_tl[i, j + 1]) * ra_y[i, j]
                             - (q[i, j] * area[i, j] + fyy[i, j] - fyy[i, j + 1]) * ra_y_tl[i, j]) / ra_y[i, j] ** 2)
            q_i[i, j] = (q[i, j] * area[i, j] + fyy[i, j] - fyy[i, j + 1]) / ra_y[i, j]
    fx, fx_tl = xppm_tlm(q_i, q_i_tl, _sub(crx, is_, ie + 1, js, je), _sub(crx_tl, is_, ie + 1, js, je), ord_ou, is_, ie, isd, ied, js, je,
                         jsd, jed, npx, npy, gs["dxa"])
    copy_corners_tlm(q, q_tl, npx, npy, 1)
    fx2, fx2_tl = xppm_tlm(q, q_tl, crx, crx_tl, ord_in, is_, ie, isd, ied, jsd, jed, jsd, jed, npx, npy, gs["dxa"])
    for j in range(jsd, jed + 1):
        for i in range(is_, ie + 2):
            fx1_tl[i] = xfx_tl[i, j] * fx2[i, j] + xfx[i, j] * fx2_tl[i, j]
            fx1[i] = xfx[i, j] * fx2[i, j]
        for i in range(is_, ie + 1):
            q_j_tl[i, j] = (((area[i, j] * q_tl[i, j] + fx1_tl[i] - fx1_tl[i + 1]) * ra_x[i, j]
                             - (q[i, j] * area[i, j] + fx1[i] - fx1[i + 1]) * ra_x_tl[i, j]) / ra_x[i, j] ** 2)
            q_j[i, j] = (q[i, j] * area[i, j] + fx1[i] - fx1[i + 1]) / ra_x[i, j]
    fy, fy_tl = yppm_tlm(q_j, q_j_tl, cry, cry_tl, ord_ou, is_, ie, isd, ied, js, je, jsd, jed, npx, npy, gs["dya"])
    if mfx is not None and mfy is not None:
        for j in range(js, je + 1):
            for i in range(is_, ie + 2):
                fx_tl[i, j] = 0.5 * ((fx_tl[i, j] + fx2_tl[i, j]) * mfx[i, j] + (fx[i, j] + fx2[i, j]) * mfx_tl[i, j])
                fx[i, j] = 0.5 * (fx[i, j] + fx2[i, j]) * mfx[i, j]
        for j in range(js, je + 2):
            for i in range(is_, ie + 1):
                fy_tl[i, j] = 0.5 * ((fy_tl[i, j] + fy2_tl[i, j]) * mfy[i, j] + (fy[i, j] + fy2[i, j]) * mfy_tl[i, j])
                fy[i, j] = 0.5 * (fy[i, j] + fy2[i, j]) * mfy[i, j]
        if nord is not None and damp_c is not None and mass is not None:
            if damp_c > 1.e-4:
                pwx1 = damp_c * gs["da_min"]
                pwy1 = nord + 1
                damp = pwx1 ** pwy1
                deln_flux_tlm(nord, is_, ie, js, je, npx, npy, damp, q, q_tl, fx, fx_tl, fy, fy_tl, gs, bd, mass, mass_tl)
    else:
        for j in range(js, je + 1):
            for i in range(is_, ie + 2):
                fx_tl[i, j] = 0.5 * ((fx_tl[i, j] + fx2_tl[i, j]) * xfx[i, j] + (fx[i, j] + fx2[i, j]) * xfx_tl[i, j])
                fx[i, j] = 0.5 * (fx[i, j] + fx2[i, j]) * xfx[i, j]
        for j in range(js, je + 2):
            for i in range(is_, ie + 1):
                fy_tl[i, j] = 0.5 * ((fy_tl[i, j] + fy2_tl[i, j]) * yfx[i, j] + (fy[i, j] + fy2[i, j]) * yfx_tl[i, j])
                fy[i, j] = 0.5 * (fy[i, j] + fy2[i, j]) * yfx[i, j]
        if nord is not None and damp_c is not None:
            if damp_c > 1.e-4:
                pwx1 = damp_c * gs["da_min"]
                pwy1 = nord + 1
                damp = pwx1 ** pwy1
                deln_flux_tlm(nord, is_, ie, js, je, npx, npy, damp, q, q_tl, fx, fx_tl, fy, fy_tl, gs, bd)
    return fx, fx_tl, fy, fy_tl
