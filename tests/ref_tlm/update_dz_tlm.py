"""UPDATE_DZ_D_TLM (model_tlmadm/nh_utils_tlm.F90:381-588) and DEL6_VT_FLUX_TLM (model_tlmadm/sw_core_tlm.F90:3621-3728), transliterated
for a whole cube tile, hord == hord_pert (the branch without the nonlinear-scheme replay)."""
from . import F
from .edge_profile_tlm import edge_profile_tlm
from .fv_tp_2d_tlm import fv_tp_2d_tlm, copy_corners_tlm

dz_min = 2.       # nh_utils_tlm.F90:36


def del6_vt_flux_tlm(nord, npx, npy, damp, q, q_tl, d2, d2_tl, fx2, fx2_tl, fy2, fy2_tl, gs, bd):
    is_, ie, js, je = bd.is_, bd.ie, bd.js, bd.je
    del6_u, del6_v, rarea = gs["del6_u"], gs["del6_v"], gs["rarea"]
    i1 = is_ - 1 - nord
    i2 = ie + 1 + nord
    j1 = js - 1 - nord
    j2 = je + 1 + nord
    for j in range(j1, j2 + 1):
        for i in range(i1, i2 + 1):
            d2_tl[i, j] = damp * q_tl[i, j]
            d2[i, j] = damp * q[i, j]
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
                    d2[i, j] = (fx2[i, j] - fx2[i + 1, j] + (fy2[i, j] - fy2[i, j + 1])) * rarea[i, j]
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


def _lev(a, k, i0, i1, j0, j1):
    """the array section a(i0:i1, j0:j1, k) as a new 2-d F"""
    s = F((i0, i1), (j0, j1))
    for j in range(j0, j1 + 1):
        for i in range(i0, i1 + 1):
            s[i, j] = a[i, j, k]
    return s


def update_dz_d_tlm(ndif, damp, hord, is_, ie, js, je, km, ng, npx, npy, dp0, zs, zh, zh_tl, crx, crx_tl, cry, cry_tl, xfx, xfx_tl,
                    yfx, yfx_tl, rdt, gs, bd):
    """zh, zh_tl: F((isd,ied),(jsd,jed),(1,km+1)) updated in place on (is:ie, js:je); returns ws, ws_tl F((is,ie),(js,je)).
    ndif, damp: python lists of length km + 1 (last entry overwritten as in the source)."""
    area, rarea = gs["area"], gs["rarea"]
    isd = is_ - ng; ied = ie + ng; jsd = js - ng; jed = je + ng
    damp[km] = damp[km - 1]
    ndif[km] = ndif[km - 1]
    crx_adv = F((is_, ie + 1), (jsd, jed), (1, km + 1)); crx_adv_tl = F((is_, ie + 1), (jsd, jed), (1, km + 1))
    xfx_adv = F((is_, ie + 1), (jsd, jed), (1, km + 1)); xfx_adv_tl = F((is_, ie + 1), (jsd, jed), (1, km + 1))
    cry_adv = F((isd, ied), (js, je + 1), (1, km + 1)); cry_adv_tl = F((isd, ied), (js, je + 1), (1, km + 1))
    yfx_adv = F((isd, ied), (js, je + 1), (1, km + 1)); yfx_adv_tl = F((isd, ied), (js, je + 1), (1, km + 1))
    ws = F((is_, ie), (js, je)); ws_tl = F((is_, ie), (js, je))
    for j in range(jsd, jed + 1):
        edge_profile_tlm(crx, crx_tl, xfx, xfx_tl, crx_adv, crx_adv_tl, xfx_adv, xfx_adv_tl, is_, ie + 1, jsd, jed, j, km, dp0, False, 0)
        if j <= je + 1 and j >= js:
            edge_profile_tlm(cry, cry_tl, yfx, yfx_tl, cry_adv, cry_adv_tl, yfx_adv, yfx_adv_tl, isd, ied, js, je + 1, j, km, dp0, False, 0)
    for k in range(1, km + 2):
        ra_x = F((is_, ie), (jsd, jed)); ra_x_tl = F((is_, ie), (jsd, jed))
        ra_y = F((isd, ied), (js, je)); ra_y_tl = F((isd, ied), (js, je))
        for j in range(jsd, jed + 1):
            for i in range(is_, ie + 1):
                ra_x_tl[i, j] = xfx_adv_tl[i, j, k] - xfx_adv_tl[i + 1, j, k]
                ra_x[i, j] = area[i, j] + (xfx_adv[i, j, k] - xfx_adv[i + 1, j, k])
        for j in range(js, je + 1):
            for i in range(isd, ied + 1):
                ra_y_tl[i, j] = yfx_adv_tl[i, j, k] - yfx_adv_tl[i, j + 1, k]
                ra_y[i, j] = area[i, j] + (yfx_adv[i, j, k] - yfx_adv[i, j + 1, k])
        sec = lambda a, xs: _lev(a, k, *xs)
        X = (is_, ie + 1, jsd, jed); Y = (isd, ied, js, je + 1)
        if damp[k - 1] > 1.e-5:
            z2 = F((isd, ied), (jsd, jed)); z2_tl = F((isd, ied), (jsd, jed))
            for j in range(jsd, jed + 1):
                for i in range(isd, ied + 1):
                    z2_tl[i, j] = zh_tl[i, j, k]
                    z2[i, j] = zh[i, j, k]
            fx, fx_tl, fy, fy_tl = fv_tp_2d_tlm(z2, z2_tl, sec(crx_adv, X), sec(crx_adv_tl, X), sec(cry_adv, Y), sec(cry_adv_tl, Y), npx, npy,
                                                hord, sec(xfx_adv, X), sec(xfx_adv_tl, X), sec(yfx_adv, Y), sec(yfx_adv_tl, Y), gs, bd,
                                                ra_x, ra_x_tl, ra_y, ra_y_tl)
            wk2 = F((isd, ied), (jsd, jed)); wk2_tl = F((isd, ied), (jsd, jed))
            fx2 = F((isd, ied + 1), (jsd, jed)); fx2_tl = F((isd, ied + 1), (jsd, jed))
            fy2 = F((isd, ied), (jsd, jed + 1)); fy2_tl = F((isd, ied), (jsd, jed + 1))
            del6_vt_flux_tlm(ndif[k - 1], npx, npy, damp[k - 1], z2, z2_tl, wk2, wk2_tl, fx2, fx2_tl, fy2, fy2_tl, gs, bd)
            for j in range(js, je + 1):
                for i in range(is_, ie + 1):
                    zh_tl[i, j, k] = (((area[i, j] * z2_tl[i, j] + fx_tl[i, j] - fx_tl[i + 1, j] + fy_tl[i, j] - fy_tl[i, j + 1])
                                       * (ra_x[i, j] + ra_y[i, j] - area[i, j])
                                       - (z2[i, j] * area[i, j] + (fx[i, j] - fx[i + 1, j]) + (fy[i, j] - fy[i, j + 1]))
                                       * (ra_x_tl[i, j] + ra_y_tl[i, j])) / (ra_x[i, j] + ra_y[i, j] - area[i, j]) ** 2
                                      + rarea[i, j] * (fx2_tl[i, j] - fx2_tl[i + 1, j] + fy2_tl[i, j] - fy2_tl[i, j + 1]))
                    zh[i, j, k] = ((z2[i, j] * area[i, j] + (fx[i, j] - fx[i + 1, j]) + (fy[i, j] - fy[i, j + 1]))
                                   / (ra_x[i, j] + ra_y[i, j] - area[i, j])
                                   + (fx2[i, j] - fx2[i + 1, j] + (fy2[i, j] - fy2[i, j + 1])) * rarea[i, j])
        else:
            zk = _lev(zh, k, isd, ied, jsd, jed); zk_tl = _lev(zh_tl, k, isd, ied, jsd, jed)
            fx, fx_tl, fy, fy_tl = fv_tp_2d_tlm(zk, zk_tl, sec(crx_adv, X), sec(crx_adv_tl, X), sec(cry_adv, Y), sec(cry_adv_tl, Y), npx, npy,
                                                hord, sec(xfx_adv, X), sec(xfx_adv_tl, X), sec(yfx_adv, Y), sec(yfx_adv_tl, Y), gs, bd,
                                                ra_x, ra_x_tl, ra_y, ra_y_tl)
            for j in range(jsd, jed + 1):       # zh(isd:ied, jsd:jed, k) is the actual argument: the corner copies land in zh itself
                for i in range(isd, ied + 1):
                    zh[i, j, k] = zk[i, j]
                    zh_tl[i, j, k] = zk_tl[i, j]
            for j in range(js, je + 1):
                for i in range(is_, ie + 1):
                    zh_tl[i, j, k] = (((area[i, j] * zh_tl[i, j, k] + fx_tl[i, j] - fx_tl[i + 1, j] + fy_tl[i, j] - fy_tl[i, j + 1])
                                       * (ra_x[i, j] + ra_y[i, j] - area[i, j])
                                       - (zh[i, j, k] * area[i, j] + (fx[i, j] - fx[i + 1, j]) + (fy[i, j] - fy[i, j + 1]))
                                       * (ra_x_tl[i, j] + ra_y_tl[i, j])) / (ra_x[i, j] + ra_y[i, j] - area[i, j]) ** 2)
                    zh[i, j, k] = ((zh[i, j, k] * area[i, j] + (fx[i, j] - fx[i + 1, j]) + (fy[i, j] - fy[i, j + 1]))
                                   / (ra_x[i, j] + ra_y[i, j] - area[i, j]))
    for j in range(js, je + 1):
        for i in range(is_, ie + 1):
            ws_tl[i, j] = -(rdt * zh_tl[i, j, km + 1])
            ws[i, j] = (zs[i, j] - zh[i, j, km + 1]) * rdt
        for k in range(km, 0, -1):
            for i in range(is_, ie + 1):
                if zh[i, j, k] < zh[i, j, k + 1] + dz_min:
                    zh_tl[i, j, k] = zh_tl[i, j, k + 1]
                    zh[i, j, k] = zh[i, j, k + 1] + dz_min
    return ws, ws_tl


def fill_4corners_tlm(q, q_tl, dir_, npx, npy):
    """FILL_4CORNERS_TLM, model_tlmadm/sw_core_tlm.F90:7138-7211, all four corners present"""
    if dir_ == 1:
        q_tl[-1, 0] = q_tl[0, 2]; q[-1, 0] = q[0, 2]
        q_tl[0, 0] = q_tl[0, 1]; q[0, 0] = q[0, 1]
        q_tl[npx + 1, 0] = q_tl[npx, 2]; q[npx + 1, 0] = q[npx, 2]
        q_tl[npx, 0] = q_tl[npx, 1]; q[npx, 0] = q[npx, 1]
        q_tl[0, npy] = q_tl[0, npy - 1]; q[0, npy] = q[0, npy - 1]
        q_tl[-1, npy] = q_tl[0, npy - 2]; q[-1, npy] = q[0, npy - 2]
        q_tl[npx, npy] = q_tl[npx, npy - 1]; q[npx, npy] = q[npx, npy - 1]
        q_tl[npx + 1, npy] = q_tl[npx, npy - 2]; q[npx + 1, npy] = q[npx, npy - 2]
    elif dir_ == 2:
        q_tl[0, 0] = q_tl[1, 0]; q[0, 0] = q[1, 0]
        q_tl[0, -1] = q_tl[2, 0]; q[0, -1] = q[2, 0]
        q_tl[npx, 0] = q_tl[npx - 1, 0]; q[npx, 0] = q[npx - 1, 0]
        q_tl[npx, -1] = q_tl[npx - 2, 0]; q[npx, -1] = q[npx - 2, 0]
        q_tl[0, npy] = q_tl[1, npy]; q[0, npy] = q[1, npy]
        q_tl[0, npy + 1] = q_tl[2, npy]; q[0, npy + 1] = q[2, npy]
        q_tl[npx, npy] = q_tl[npx - 1, npy]; q[npx, npy] = q[npx - 1, npy]
        q_tl[npx, npy + 1] = q_tl[npx - 2, npy]; q[npx, npy + 1] = q[npx - 2, npy]


def update_dz_c_tlm(is_, ie, js, je, km, ng, dt, dp0, zs, area, ut, ut_tl, vt, vt_tl, gz, gz_tl, npx, npy):
    """UPDATE_DZ_C_TLM, model_tlmadm/nh_utils_tlm.F90:51-233.  gz, gz_tl: F((isd,ied),(jsd,jed),(1,km+1)) updated in place on
    (is-1:ie+1, js-1:je+1); returns ws, ws_tl F((isd,ied),(jsd,jed))."""
    isd = is_ - ng; ied = ie + ng; jsd = js - ng; jed = je + ng
    ws = F((isd, ied), (jsd, jed)); ws_tl = F((isd, ied), (jsd, jed))
    gz2 = F((isd, ied), (jsd, jed)); gz2_tl = F((isd, ied), (jsd, jed))
    xfx = F((is_ - 1, ie + 2), (js - 1, je + 1)); fx = F((is_ - 1, ie + 2), (js - 1, je + 1))
    xfx_tl = F((is_ - 1, ie + 2), (js - 1, je + 1)); fx_tl = F((is_ - 1, ie + 2), (js - 1, je + 1))
    yfx = F((is_ - 1, ie + 1), (js - 1, je + 2)); fy = F((is_ - 1, ie + 1), (js - 1, je + 2))
    yfx_tl = F((is_ - 1, ie + 1), (js - 1, je + 2)); fy_tl = F((is_ - 1, ie + 1), (js - 1, je + 2))
    rdt = 1. / dt
    top_ratio = dp0[1] / (dp0[1] + dp0[2])
    bot_ratio = dp0[km] / (dp0[km - 1] + dp0[km])
    is1 = is_ - 1; js1 = js - 1; ie1 = ie + 1; je1 = je + 1; ie2 = ie + 2; je2 = je + 2
    for k in range(1, km + 2):
        if k == 1:
            for j in range(js1, je1 + 1):
                for i in range(is1, ie2 + 1):
                    xfx_tl[i, j] = ut_tl[i, j, 1] + top_ratio * (ut_tl[i, j, 1] - ut_tl[i, j, 2])
                    xfx[i, j] = ut[i, j, 1] + (ut[i, j, 1] - ut[i, j, 2]) * top_ratio
            for j in range(js1, je2 + 1):
                for i in range(is1, ie1 + 1):
                    yfx_tl[i, j] = vt_tl[i, j, 1] + top_ratio * (vt_tl[i, j, 1] - vt_tl[i, j, 2])
                    yfx[i, j] = vt[i, j, 1] + (vt[i, j, 1] - vt[i, j, 2]) * top_ratio
        elif k == km + 1:
            for j in range(js1, je1 + 1):
                for i in range(is1, ie2 + 1):
                    xfx_tl[i, j] = ut_tl[i, j, km] + bot_ratio * (ut_tl[i, j, km] - ut_tl[i, j, km - 1])
                    xfx[i, j] = ut[i, j, km] + (ut[i, j, km] - ut[i, j, km - 1]) * bot_ratio
            for j in range(js1, je2 + 1):
                for i in range(is1, ie1 + 1):
                    yfx_tl[i, j] = vt_tl[i, j, km] + bot_ratio * (vt_tl[i, j, km] - vt_tl[i, j, km - 1])
                    yfx[i, j] = vt[i, j, km] + (vt[i, j, km] - vt[i, j, km - 1]) * bot_ratio
        else:
            int_ratio = 1. / (dp0[k - 1] + dp0[k])
            for j in range(js1, je1 + 1):
                for i in range(is1, ie2 + 1):
                    xfx_tl[i, j] = int_ratio * (dp0[k] * ut_tl[i, j, k - 1] + dp0[k - 1] * ut_tl[i, j, k])
                    xfx[i, j] = (dp0[k] * ut[i, j, k - 1] + dp0[k - 1] * ut[i, j, k]) * int_ratio
            for j in range(js1, je2 + 1):
                for i in range(is1, ie1 + 1):
                    yfx_tl[i, j] = int_ratio * (dp0[k] * vt_tl[i, j, k - 1] + dp0[k - 1] * vt_tl[i, j, k])
                    yfx[i, j] = (dp0[k] * vt[i, j, k - 1] + dp0[k - 1] * vt[i, j, k]) * int_ratio
        for j in range(js - ng, je + ng + 1):
            for i in range(is_ - ng, ie + ng + 1):
                gz2_tl[i, j] = gz_tl[i, j, k]
                gz2[i, j] = gz[i, j, k]
        fill_4corners_tlm(gz2, gz2_tl, 1, npx, npy)
        for j in range(js1, je1 + 1):
            for i in range(is1, ie2 + 1):
                if xfx[i, j] > 0.:
                    fx_tl[i, j] = gz2_tl[i - 1, j]
                    fx[i, j] = gz2[i - 1, j]
                else:
                    fx_tl[i, j] = gz2_tl[i, j]
                    fx[i, j] = gz2[i, j]
                fx_tl[i, j] = xfx_tl[i, j] * fx[i, j] + xfx[i, j] * fx_tl[i, j]
                fx[i, j] = xfx[i, j] * fx[i, j]
        fill_4corners_tlm(gz2, gz2_tl, 2, npx, npy)
        for j in range(js1, je2 + 1):
            for i in range(is1, ie1 + 1):
                if yfx[i, j] > 0.:
                    fy_tl[i, j] = gz2_tl[i, j - 1]
                    fy[i, j] = gz2[i, j - 1]
                else:
                    fy_tl[i, j] = gz2_tl[i, j]
                    fy[i, j] = gz2[i, j]
                fy_tl[i, j] = yfx_tl[i, j] * fy[i, j] + yfx[i, j] * fy_tl[i, j]
                fy[i, j] = yfx[i, j] * fy[i, j]
        for j in range(js1, je1 + 1):
            for i in range(is1, ie1 + 1):
                den = area[i, j] + (xfx[i, j] - xfx[i + 1, j]) + (yfx[i, j] - yfx[i, j + 1])
                gz_tl[i, j, k] = (((area[i, j] * gz2_tl[i, j] + fx_tl[i, j] - fx_tl[i + 1, j] + fy_tl[i, j] - fy_tl[i, j + 1]) * den
                                   - (gz2[i, j] * area[i, j] + (fx[i, j] - fx[i + 1, j]) + (fy[i, j] - fy[i, j + 1]))
                                   * (xfx_tl[i, j] - xfx_tl[i + 1, j] + yfx_tl[i, j] - yfx_tl[i, j + 1])) / den ** 2)
                gz[i, j, k] = (gz2[i, j] * area[i, j] + (fx[i, j] - fx[i + 1, j]) + (fy[i, j] - fy[i, j + 1])) / den
    for j in range(js1, je1 + 1):
        for i in range(is1, ie1 + 1):
            ws_tl[i, j] = -(rdt * gz_tl[i, j, km + 1])
            ws[i, j] = (zs[i, j] - gz[i, j, km + 1]) * rdt
        for k in range(km, 0, -1):
            for i in range(is1, ie1 + 1):
                if gz[i, j, k] < gz[i, j, k + 1] + dz_min:
                    gz_tl[i, j, k] = gz_tl[i, j, k + 1]
                    gz[i, j, k] = gz[i, j, k + 1] + dz_min
    return ws, ws_tl
