"""C_SW_TLM (model_tlmadm/sw_core_tlm.F90:87-645), transliterated for a whole cube tile (not nested, grid_type 0): the C-grid half step of
the shallow-water core.  FILL2_4CORNERS_TLM (:6953-7077) is FILL_4CORNERS_TLM's statements applied to two fields."""
from . import F
from .d2a2c_vect_tlm import d2a2c_vect_tlm, divergence_corner_tlm
from .update_dz_tlm import fill_4corners_tlm


def fill2_4corners_tlm(q1, q1_tl, q2, q2_tl, dir_, npx, npy):
    fill_4corners_tlm(q1, q1_tl, dir_, npx, npy)
    fill_4corners_tlm(q2, q2_tl, dir_, npx, npy)


def c_sw_tlm(delp, delp_tl, pt, pt_tl, u, u_tl, v, v_tl, w, w_tl, nord, dt2, hydrostatic, dord4, bd, gs, npx, npy):
    """delp, pt, w (and _tl) get their corner fills in place as in the source.  gs: the F arrays / callables of d2a2c_vect_tlm and
    divergence_corner_tlm plus dx, dy, sina_u, sina_v, rarea, rdxc, rdyc, fC.  Returns a dict of F arrays (delpc, ptc, wc, uc, vc, ua,
    va, ut, vt, divg_d and their _tl)."""
    is_, ie, js, je = bd.is_, bd.ie, bd.js, bd.je
    isd, ied, jsd, jed = bd.isd, bd.ied, bd.jsd, bd.jed
    sin_sg, cos_sg = gs["sin_sg"], gs["cos_sg"]
    cosa_u, cosa_v, sina_u, sina_v = gs["cosa_u"], gs["cosa_v"], gs["sina_u"], gs["sina_v"]
    dx, dy, dxc, dyc = gs["dx"], gs["dy"], gs["dxc"], gs["dyc"]
    rarea, rarea_c, rdxc, rdyc, fC = gs["rarea"], gs["rarea_c"], gs["rdxc"], gs["rdyc"], gs["fC"]
    A = lambda: F((isd, ied), (jsd, jed))
    delpc, ptc, wc, delpc_tl, ptc_tl, wc_tl = A(), A(), A(), A(), A(), A()
    B1 = ((is_ - 1, ie + 1), (js - 1, je + 1)); BX = ((is_ - 1, ie + 2), (js - 1, je + 1)); BY = ((is_ - 1, ie + 1), (js - 1, je + 2))
    vort, ke, vort_tl, ke_tl = F(*B1), F(*B1), F(*B1), F(*B1)
    fx, fx1, fx2, fx_tl, fx1_tl, fx2_tl = F(*BX), F(*BX), F(*BX), F(*BX), F(*BX), F(*BX)
    fy, fy1, fy2, fy_tl, fy1_tl, fy2_tl = F(*BY), F(*BY), F(*BY), F(*BY), F(*BY), F(*BY)
    iep1 = ie + 1
    jep1 = je + 1
    r = d2a2c_vect_tlm(u, u_tl, v, v_tl, dord4, gs, bd, npx, npy)
    ua, va, uc, vc, ut, vt = r["ua"], r["va"], r["uc"], r["vc"], r["ut"], r["vt"]
    ua_tl, va_tl, uc_tl, vc_tl, ut_tl, vt_tl = r["ua_tl"], r["va_tl"], r["uc_tl"], r["vc_tl"], r["ut_tl"], r["vt_tl"]
    divg_d = divg_d_tl = None
    if nord > 0:
        divg_d, divg_d_tl = divergence_corner_tlm(u, u_tl, v, v_tl, ua, ua_tl, va, va_tl, gs, bd, npx, npy)
    for j in range(js - 1, jep1 + 1):
        for i in range(is_ - 1, iep1 + 2):
            if ut[i, j] > 0.:
                ut_tl[i, j] = dt2 * dy[i, j] * sin_sg(i - 1, j, 3) * ut_tl[i, j]
                ut[i, j] = dt2 * ut[i, j] * dy[i, j] * sin_sg(i - 1, j, 3)
            else:
                ut_tl[i, j] = dt2 * dy[i, j] * sin_sg(i, j, 1) * ut_tl[i, j]
                ut[i, j] = dt2 * ut[i, j] * dy[i, j] * sin_sg(i, j, 1)
    for j in range(js - 1, je + 3):
        for i in range(is_ - 1, iep1 + 1):
            if vt[i, j] > 0.:
                vt_tl[i, j] = dt2 * dx[i, j] * sin_sg(i, j - 1, 4) * vt_tl[i, j]
                vt[i, j] = dt2 * vt[i, j] * dx[i, j] * sin_sg(i, j - 1, 4)
            else:
                vt_tl[i, j] = dt2 * dx[i, j] * sin_sg(i, j, 2) * vt_tl[i, j]
                vt[i, j] = dt2 * vt[i, j] * dx[i, j] * sin_sg(i, j, 2)
    # Transport delp, Xdir
    fill2_4corners_tlm(delp, delp_tl, pt, pt_tl, 1, npx, npy)
    if hydrostatic:
        for j in range(js - 1, jep1 + 1):
            for i in range(is_ - 1, ie + 3):
                if ut[i, j] > 0.:
                    fx1_tl[i, j] = delp_tl[i - 1, j]; fx1[i, j] = delp[i - 1, j]
                    fx_tl[i, j] = pt_tl[i - 1, j]; fx[i, j] = pt[i - 1, j]
                else:
                    fx1_tl[i, j] = delp_tl[i, j]; fx1[i, j] = delp[i, j]
                    fx_tl[i, j] = pt_tl[i, j]; fx[i, j] = pt[i, j]
                fx1_tl[i, j] = ut_tl[i, j] * fx1[i, j] + ut[i, j] * fx1_tl[i, j]
                fx1[i, j] = ut[i, j] * fx1[i, j]
                fx_tl[i, j] = fx1_tl[i, j] * fx[i, j] + fx1[i, j] * fx_tl[i, j]
                fx[i, j] = fx1[i, j] * fx[i, j]
    else:
        fill_4corners_tlm(w, w_tl, 1, npx, npy)
        for j in range(js - 1, je + 2):
            for i in range(is_ - 1, ie + 3):
                if ut[i, j] > 0.:
                    fx1_tl[i, j] = delp_tl[i - 1, j]; fx1[i, j] = delp[i - 1, j]
                    fx_tl[i, j] = pt_tl[i - 1, j]; fx[i, j] = pt[i - 1, j]
                    fx2_tl[i, j] = w_tl[i - 1, j]; fx2[i, j] = w[i - 1, j]
                else:
                    fx1_tl[i, j] = delp_tl[i, j]; fx1[i, j] = delp[i, j]
                    fx_tl[i, j] = pt_tl[i, j]; fx[i, j] = pt[i, j]
                    fx2_tl[i, j] = w_tl[i, j]; fx2[i, j] = w[i, j]
                fx1_tl[i, j] = ut_tl[i, j] * fx1[i, j] + ut[i, j] * fx1_tl[i, j]
                fx1[i, j] = ut[i, j] * fx1[i, j]
                fx_tl[i, j] = fx1_tl[i, j] * fx[i, j] + fx1[i, j] * fx_tl[i, j]
                fx[i, j] = fx1[i, j] * fx[i, j]
                fx2_tl[i, j] = fx1_tl[i, j] * fx2[i, j] + fx1[i, j] * fx2_tl[i, j]
                fx2[i, j] = fx1[i, j] * fx2[i, j]
    # Ydir
    fill2_4corners_tlm(delp, delp_tl, pt, pt_tl, 2, npx, npy)
    if hydrostatic:
        for j in range(js - 1, jep1 + 2):
            for i in range(is_ - 1, iep1 + 1):
                if vt[i, j] > 0.:
                    fy1_tl[i, j] = delp_tl[i, j - 1]; fy1[i, j] = delp[i, j - 1]
                    fy_tl[i, j] = pt_tl[i, j - 1]; fy[i, j] = pt[i, j - 1]
                else:
                    fy1_tl[i, j] = delp_tl[i, j]; fy1[i, j] = delp[i, j]
                    fy_tl[i, j] = pt_tl[i, j]; fy[i, j] = pt[i, j]
                fy1_tl[i, j] = vt_tl[i, j] * fy1[i, j] + vt[i, j] * fy1_tl[i, j]
                fy1[i, j] = vt[i, j] * fy1[i, j]
                fy_tl[i, j] = fy1_tl[i, j] * fy[i, j] + fy1[i, j] * fy_tl[i, j]
                fy[i, j] = fy1[i, j] * fy[i, j]
        for j in range(js - 1, jep1 + 1):
            for i in range(is_ - 1, iep1 + 1):
                delpc_tl[i, j] = delp_tl[i, j] + rarea[i, j] * (fx1_tl[i, j] - fx1_tl[i + 1, j] + fy1_tl[i, j] - fy1_tl[i, j + 1])
                delpc[i, j] = delp[i, j] + (fx1[i, j] - fx1[i + 1, j] + (fy1[i, j] - fy1[i, j + 1])) * rarea[i, j]
                ptc_tl[i, j] = (((pt_tl[i, j] * delp[i, j] + pt[i, j] * delp_tl[i, j]
                                  + rarea[i, j] * (fx_tl[i, j] - fx_tl[i + 1, j] + fy_tl[i, j] - fy_tl[i, j + 1])) * delpc[i, j]
                                 - (pt[i, j] * delp[i, j] + (fx[i, j] - fx[i + 1, j] + (fy[i, j] - fy[i, j + 1])) * rarea[i, j]) * delpc_tl[i, j])
                                / delpc[i, j] ** 2)
                ptc[i, j] = (pt[i, j] * delp[i, j] + (fx[i, j] - fx[i + 1, j] + (fy[i, j] - fy[i, j + 1])) * rarea[i, j]) / delpc[i, j]
    else:
        fill_4corners_tlm(w, w_tl, 2, npx, npy)
        for j in range(js - 1, je + 3):
            for i in range(is_ - 1, ie + 2):
                if vt[i, j] > 0.:
                    fy1_tl[i, j] = delp_tl[i, j - 1]; fy1[i, j] = delp[i, j - 1]
                    fy_tl[i, j] = pt_tl[i, j - 1]; fy[i, j] = pt[i, j - 1]
                    fy2_tl[i, j] = w_tl[i, j - 1]; fy2[i, j] = w[i, j - 1]
                else:
                    fy1_tl[i, j] = delp_tl[i, j]; fy1[i, j] = delp[i, j]
                    fy_tl[i, j] = pt_tl[i, j]; fy[i, j] = pt[i, j]
                    fy2_tl[i, j] = w_tl[i, j]; fy2[i, j] = w[i, j]
                fy1_tl[i, j] = vt_tl[i, j] * fy1[i, j] + vt[i, j] * fy1_tl[i, j]
                fy1[i, j] = vt[i, j] * fy1[i, j]
                fy_tl[i, j] = fy1_tl[i, j] * fy[i, j] + fy1[i, j] * fy_tl[i, j]
                fy[i, j] = fy1[i, j] * fy[i, j]
                fy2_tl[i, j] = fy1_tl[i, j] * fy2[i, j] + fy1[i, j] * fy2_tl[i, j]
                fy2[i, j] = fy1[i, j] * fy2[i, j]
        for j in range(js - 1, je + 2):
            for i in range(is_ - 1, ie + 2):
                delpc_tl[i, j] = delp_tl[i, j] + rarea[i, j] * (fx1_tl[i, j] - fx1_tl[i + 1, j] + fy1_tl[i, j] - fy1_tl[i, j + 1])
                delpc[i, j] = delp[i, j] + (fx1[i, j] - fx1[i + 1, j] + (fy1[i, j] - fy1[i, j + 1])) * rarea[i, j]
                ptc_tl[i, j] = (((pt_tl[i, j] * delp[i, j] + pt[i, j] * delp_tl[i, j]
                                  + rarea[i, j] * (fx_tl[i, j] - fx_tl[i + 1, j] + fy_tl[i, j] - fy_tl[i, j + 1])) * delpc[i, j]
                                 - (pt[i, j] * delp[i, j] + (fx[i, j] - fx[i + 1, j] + (fy[i, j] - fy[i, j + 1])) * rarea[i, j]) * delpc_tl[i, j])
                                / delpc[i, j] ** 2)
                ptc[i, j] = (pt[i, j] * delp[i, j] + (fx[i, j] - fx[i + 1, j] + (fy[i, j] - fy[i, j + 1])) * rarea[i, j]) / delpc[i, j]
                wc_tl[i, j] = (((w_tl[i, j] * delp[i, j] + w[i, j] * delp_tl[i, j]
                                 + rarea[i, j] * (fx2_tl[i, j] - fx2_tl[i + 1, j] + fy2_tl[i, j] - fy2_tl[i, j + 1])) * delpc[i, j]
                                - (w[i, j] * delp[i, j] + (fx2[i, j] - fx2[i + 1, j] + (fy2[i, j] - fy2[i, j + 1])) * rarea[i, j]) * delpc_tl[i, j])
                               / delpc[i, j] ** 2)
                wc[i, j] = (w[i, j] * delp[i, j] + (fx2[i, j] - fx2[i + 1, j] + (fy2[i, j] - fy2[i, j + 1])) * rarea[i, j]) / delpc[i, j]
    # Compute KE
    for j in range(js - 1, jep1 + 1):
        for i in range(is_ - 1, iep1 + 1):
            if ua[i, j] > 0.:
                if i == 1:
                    ke_tl[1, j] = sin_sg(1, j, 1) * uc_tl[1, j] + cos_sg(1, j, 1) * v_tl[1, j]
                    ke[1, j] = uc[1, j] * sin_sg(1, j, 1) + v[1, j] * cos_sg(1, j, 1)
                elif i == npx:
                    ke_tl[i, j] = sin_sg(npx, j, 1) * uc_tl[npx, j] + cos_sg(npx, j, 1) * v_tl[npx, j]
                    ke[i, j] = uc[npx, j] * sin_sg(npx, j, 1) + v[npx, j] * cos_sg(npx, j, 1)
                else:
                    ke_tl[i, j] = uc_tl[i, j]
                    ke[i, j] = uc[i, j]
            elif i == 0:
                ke_tl[0, j] = sin_sg(0, j, 3) * uc_tl[1, j] + cos_sg(0, j, 3) * v_tl[1, j]
                ke[0, j] = uc[1, j] * sin_sg(0, j, 3) + v[1, j] * cos_sg(0, j, 3)
            elif i == npx - 1:
                ke_tl[i, j] = sin_sg(npx - 1, j, 3) * uc_tl[npx, j] + cos_sg(npx - 1, j, 3) * v_tl[npx, j]
                ke[i, j] = uc[npx, j] * sin_sg(npx - 1, j, 3) + v[npx, j] * cos_sg(npx - 1, j, 3)
            else:
                ke_tl[i, j] = uc_tl[i + 1, j]
                ke[i, j] = uc[i + 1, j]
    for j in range(js - 1, jep1 + 1):
        for i in range(is_ - 1, iep1 + 1):
            if va[i, j] > 0.:
                if j == 1:
                    vort_tl[i, 1] = sin_sg(i, 1, 2) * vc_tl[i, 1] + cos_sg(i, 1, 2) * u_tl[i, 1]
                    vort[i, 1] = vc[i, 1] * sin_sg(i, 1, 2) + u[i, 1] * cos_sg(i, 1, 2)
                elif j == npy:
                    vort_tl[i, j] = sin_sg(i, npy, 2) * vc_tl[i, npy] + cos_sg(i, npy, 2) * u_tl[i, npy]
                    vort[i, j] = vc[i, npy] * sin_sg(i, npy, 2) + u[i, npy] * cos_sg(i, npy, 2)
                else:
                    vort_tl[i, j] = vc_tl[i, j]
                    vort[i, j] = vc[i, j]
            elif j == 0:
                vort_tl[i, 0] = sin_sg(i, 0, 4) * vc_tl[i, 1] + cos_sg(i, 0, 4) * u_tl[i, 1]
                vort[i, 0] = vc[i, 1] * sin_sg(i, 0, 4) + u[i, 1] * cos_sg(i, 0, 4)
            elif j == npy - 1:
                vort_tl[i, j] = sin_sg(i, npy - 1, 4) * vc_tl[i, npy] + cos_sg(i, npy - 1, 4) * u_tl[i, npy]
                vort[i, j] = vc[i, npy] * sin_sg(i, npy - 1, 4) + u[i, npy] * cos_sg(i, npy - 1, 4)
            else:
                vort_tl[i, j] = vc_tl[i, j + 1]
                vort[i, j] = vc[i, j + 1]
    dt4 = 0.5 * dt2
    for j in range(js - 1, jep1 + 1):
        for i in range(is_ - 1, iep1 + 1):
            ke_tl[i, j] = dt4 * (ua_tl[i, j] * ke[i, j] + ua[i, j] * ke_tl[i, j] + va_tl[i, j] * vort[i, j] + va[i, j] * vort_tl[i, j])
            ke[i, j] = dt4 * (ua[i, j] * ke[i, j] + va[i, j] * vort[i, j])
    # circulation on the C grid
    for j in range(js - 1, je + 2):
        for i in range(is_, ie + 2):
            fx_tl[i, j] = dxc[i, j] * uc_tl[i, j]
            fx[i, j] = uc[i, j] * dxc[i, j]
    for j in range(js, je + 2):
        for i in range(is_ - 1, ie + 2):
            fy_tl[i, j] = dyc[i, j] * vc_tl[i, j]
            fy[i, j] = vc[i, j] * dyc[i, j]
    for j in range(js, je + 2):
        for i in range(is_, ie + 2):
            vort_tl[i, j] = fx_tl[i, j - 1] - fx_tl[i, j] + fy_tl[i, j] - fy_tl[i - 1, j]
            vort[i, j] = fx[i, j - 1] - fx[i, j] + (fy[i, j] - fy[i - 1, j])
    vort_tl[1, 1] = vort_tl[1, 1] + fy_tl[0, 1]
    vort[1, 1] = vort[1, 1] + fy[0, 1]
    vort_tl[npx, 1] = vort_tl[npx, 1] - fy_tl[npx, 1]
    vort[npx, 1] = vort[npx, 1] - fy[npx, 1]
    vort_tl[npx, npy] = vort_tl[npx, npy] - fy_tl[npx, npy]
    vort[npx, npy] = vort[npx, npy] - fy[npx, npy]
    vort_tl[1, npy] = vort_tl[1, npy] + fy_tl[0, npy]
    vort[1, npy] = vort[1, npy] + fy[0, npy]
    for j in range(js, je + 2):
        for i in range(is_, ie + 2):
            vort_tl[i, j] = rarea_c[i, j] * vort_tl[i, j]
            vort[i, j] = fC[i, j] + rarea_c[i, j] * vort[i, j]
    # transport of absolute vorticity
    for j in range(js, je + 1):
        for i in range(is_, iep1 + 1):
            if i == 1 or i == npx:
                fy1_tl[i, j] = dt2 * v_tl[i, j]
                fy1[i, j] = dt2 * v[i, j]
            else:
                fy1_tl[i, j] = dt2 * (v_tl[i, j] - cosa_u[i, j] * uc_tl[i, j]) / sina_u[i, j]
                fy1[i, j] = dt2 * (v[i, j] - uc[i, j] * cosa_u[i, j]) / sina_u[i, j]
            if fy1[i, j] > 0.:
                fy_tl[i, j] = vort_tl[i, j]
                fy[i, j] = vort[i, j]
            else:
                fy_tl[i, j] = vort_tl[i, j + 1]
                fy[i, j] = vort[i, j + 1]
    for j in range(js, jep1 + 1):
        if j == 1 or j == npy:
            for i in range(is_, ie + 1):
                fx1_tl[i, j] = dt2 * u_tl[i, j]
                fx1[i, j] = dt2 * u[i, j]
                if fx1[i, j] > 0.:
                    fx_tl[i, j] = vort_tl[i, j]
                    fx[i, j] = vort[i, j]
                else:
                    fx_tl[i, j] = vort_tl[i + 1, j]
                    fx[i, j] = vort[i + 1, j]
        else:
            for i in range(is_, ie + 1):
                fx1_tl[i, j] = dt2 * (u_tl[i, j] - cosa_v[i, j] * vc_tl[i, j]) / sina_v[i, j]
                fx1[i, j] = dt2 * (u[i, j] - vc[i, j] * cosa_v[i, j]) / sina_v[i, j]
                if fx1[i, j] > 0.:
                    fx_tl[i, j] = vort_tl[i, j]
                    fx[i, j] = vort[i, j]
                else:
                    fx_tl[i, j] = vort_tl[i + 1, j]
                    fx[i, j] = vort[i + 1, j]
    for j in range(js, je + 1):
        for i in range(is_, iep1 + 1):
            uc_tl[i, j] = uc_tl[i, j] + fy1_tl[i, j] * fy[i, j] + fy1[i, j] * fy_tl[i, j] + rdxc[i, j] * (ke_tl[i - 1, j] - ke_tl[i, j])
            uc[i, j] = uc[i, j] + fy1[i, j] * fy[i, j] + rdxc[i, j] * (ke[i - 1, j] - ke[i, j])
    for j in range(js, jep1 + 1):
        for i in range(is_, ie + 1):
            vc_tl[i, j] = vc_tl[i, j] - fx1_tl[i, j] * fx[i, j] - fx1[i, j] * fx_tl[i, j] + rdyc[i, j] * (ke_tl[i, j - 1] - ke_tl[i, j])
            vc[i, j] = vc[i, j] - fx1[i, j] * fx[i, j] + rdyc[i, j] * (ke[i, j - 1] - ke[i, j])
    return dict(delpc=delpc, ptc=ptc, wc=wc, uc=uc, vc=vc, ua=ua, va=va, ut=ut, vt=vt, divg_d=divg_d,
                delpc_tl=delpc_tl, ptc_tl=ptc_tl, wc_tl=wc_tl, uc_tl=uc_tl, vc_tl=vc_tl, ua_tl=ua_tl, va_tl=va_tl, ut_tl=ut_tl, vt_tl=vt_tl,
                divg_d_tl=divg_d_tl)
