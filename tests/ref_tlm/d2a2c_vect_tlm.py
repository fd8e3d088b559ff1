"""D2A2C_VECT_TLM (model_tlmadm/sw_core_tlm.F90:5874-6395, with EDGE_INTERPOLATE4_TLM :6810) and DIVERGENCE_CORNER_TLM (:3805-3966),
transliterated for a whole cube tile (not nested, grid_type 0, all four corners present).  sin_sg / cos_sg are callables (i, j, n)."""
from . import F

a1, a2 = 0.5625, -0.0625                       # sw_core_tlm.F90:56-57
c1, c2, c3 = -2. / 14., 11. / 14., 5. / 14.    # :60-62
big_number = 1.e30                             # :46


def edge_interpolate4_tlm(ua, ua_tl, dxa):
    """ua, ua_tl, dxa: 4-sequences.  Returns (tangent, value)"""
    t1 = dxa[0] + dxa[1]
    t2 = dxa[2] + dxa[3]
    tl = 0.5 * (((t1 + dxa[1]) * ua_tl[1] - dxa[1] * ua_tl[0]) / t1 + ((t2 + dxa[2]) * ua_tl[2] - dxa[2] * ua_tl[3]) / t2)
    val = 0.5 * (((t1 + dxa[1]) * ua[1] - dxa[1] * ua[0]) / t1 + ((t2 + dxa[2]) * ua[2] - dxa[2] * ua[3]) / t2)
    return tl, val


def d2a2c_vect_tlm(u, u_tl, v, v_tl, dord4, gs, bd, npx, npy):
    """gs: dict of F arrays cosa_u, cosa_v, cosa_s, rsin_u, rsin_v, rsin2, dxa, dya and the callable sin_sg(i, j, n).
    Returns dict of F arrays ua, va, uc, vc, ut, vt and their _tl."""
    is_, ie, js, je = bd.is_, bd.ie, bd.js, bd.je
    isd, ied, jsd, jed = bd.isd, bd.ied, bd.jsd, bd.jed
    sin_sg = gs["sin_sg"]
    cosa_u, cosa_v, cosa_s = gs["cosa_u"], gs["cosa_v"], gs["cosa_s"]
    rsin_u, rsin_v, rsin2 = gs["rsin_u"], gs["rsin_v"], gs["rsin2"]
    dxa, dya = gs["dxa"], gs["dya"]
    A = lambda: F((isd, ied), (jsd, jed))
    uc = F((isd, ied + 1), (jsd, jed)); uc_tl = F((isd, ied + 1), (jsd, jed))
    vc = F((isd, ied), (jsd, jed + 1)); vc_tl = F((isd, ied), (jsd, jed + 1))
    ua, va, ut, vt, ua_tl, va_tl, ut_tl, vt_tl = A(), A(), A(), A(), A(), A(), A(), A()
    utmp, vtmp, utmp_tl, vtmp_tl = A(), A(), A(), A()
    id_ = 1 if dord4 else 0
    npt = 4
    utmp.fill(big_number)
    vtmp.fill(big_number)
    max1 = js - 1 if npt < js - 1 else npt
    min1 = je + 1 if npy - npt > je + 1 else npy - npt
    for j in range(max1, min1 + 1):
        max2 = isd if npt < isd else npt
        min2 = ied if npx - npt > ied else npx - npt
        for i in range(max2, min2 + 1):
            utmp_tl[i, j] = a2 * (u_tl[i, j - 1] + u_tl[i, j + 2]) + a1 * (u_tl[i, j] + u_tl[i, j + 1])
            utmp[i, j] = a2 * (u[i, j - 1] + u[i, j + 2]) + a1 * (u[i, j] + u[i, j + 1])
    max3 = jsd if npt < jsd else npt
    min3 = jed if npy - npt > jed else npy - npt
    for j in range(max3, min3 + 1):
        max4 = is_ - 1 if npt < is_ - 1 else npt
        min4 = ie + 1 if npx - npt > ie + 1 else npx - npt
        for i in range(max4, min4 + 1):
            vtmp_tl[i, j] = a2 * (v_tl[i - 1, j] + v_tl[i + 2, j]) + a1 * (v_tl[i, j] + v_tl[i + 1, j])
            vtmp[i, j] = a2 * (v[i - 1, j] + v[i + 2, j]) + a1 * (v[i, j] + v[i + 1, j])

    def two_point(i, j):
        utmp_tl[i, j] = 0.5 * (u_tl[i, j] + u_tl[i, j + 1])
        utmp[i, j] = 0.5 * (u[i, j] + u[i, j + 1])
        vtmp_tl[i, j] = 0.5 * (v_tl[i, j] + v_tl[i + 1, j])
        vtmp[i, j] = 0.5 * (v[i, j] + v[i + 1, j])
    if js == 1 or jsd < npt:
        for j in range(jsd, npt):
            for i in range(isd, ied + 1):
                two_point(i, j)
    if je + 1 == npy or jed >= npy - npt:
        for j in range(npy - npt + 1, jed + 1):
            for i in range(isd, ied + 1):
                two_point(i, j)
    if is_ == 1 or isd < npt:
        max5 = jsd if npt < jsd else npt
        min5 = jed if npy - npt > jed else npy - npt
        for j in range(max5, min5 + 1):
            for i in range(isd, npt):
                two_point(i, j)
    if ie + 1 == npx or ied >= npx - npt:
        max6 = jsd if npt < jsd else npt
        min6 = jed if npy - npt > jed else npy - npt
        for j in range(max6, min6 + 1):
            for i in range(npx - npt + 1, ied + 1):
                two_point(i, j)
    for j in range(js - 1 - id_, je + 1 + id_ + 1):
        for i in range(is_ - 1 - id_, ie + 1 + id_ + 1):
            ua_tl[i, j] = rsin2[i, j] * (utmp_tl[i, j] - cosa_s[i, j] * vtmp_tl[i, j])
            ua[i, j] = (utmp[i, j] - vtmp[i, j] * cosa_s[i, j]) * rsin2[i, j]
            va_tl[i, j] = rsin2[i, j] * (vtmp_tl[i, j] - cosa_s[i, j] * utmp_tl[i, j])
            va[i, j] = (vtmp[i, j] - utmp[i, j] * cosa_s[i, j]) * rsin2[i, j]
    # A -> C, Xdir: fix the edges
    for i in range(-2, 1):
        utmp_tl[i, 0] = -vtmp_tl[0, 1 - i]
        utmp[i, 0] = -vtmp[0, 1 - i]
    for i in range(0, 3):
        utmp_tl[npx + i, 0] = vtmp_tl[npx, i + 1]
        utmp[npx + i, 0] = vtmp[npx, i + 1]
    for i in range(0, 3):
        utmp_tl[npx + i, npy] = -vtmp_tl[npx, je - i]
        utmp[npx + i, npy] = -vtmp[npx, je - i]
    for i in range(-2, 1):
        utmp_tl[i, npy] = vtmp_tl[0, je + i]
        utmp[i, npy] = vtmp[0, je + i]
    ifirst = is_ - 1 if 3 < is_ - 1 else 3
    ilast = ie + 2 if npx - 2 > ie + 2 else npx - 2
    for j in range(js - 1, je + 2):
        for i in range(ifirst, ilast + 1):
            uc_tl[i, j] = a2 * (utmp_tl[i - 2, j] + utmp_tl[i + 1, j]) + a1 * (utmp_tl[i - 1, j] + utmp_tl[i, j])
            uc[i, j] = a2 * (utmp[i - 2, j] + utmp[i + 1, j]) + a1 * (utmp[i - 1, j] + utmp[i, j])
            ut_tl[i, j] = rsin_u[i, j] * (uc_tl[i, j] - cosa_u[i, j] * v_tl[i, j])
            ut[i, j] = (uc[i, j] - v[i, j] * cosa_u[i, j]) * rsin_u[i, j]
    ua_tl[-1, 0] = -va_tl[0, 2]; ua[-1, 0] = -va[0, 2]
    ua_tl[0, 0] = -va_tl[0, 1]; ua[0, 0] = -va[0, 1]
    ua_tl[npx, 0] = va_tl[npx, 1]; ua[npx, 0] = va[npx, 1]
    ua_tl[npx + 1, 0] = va_tl[npx, 2]; ua[npx + 1, 0] = va[npx, 2]
    ua_tl[npx, npy] = -va_tl[npx, npy - 1]; ua[npx, npy] = -va[npx, npy - 1]
    ua_tl[npx + 1, npy] = -va_tl[npx, npy - 2]; ua[npx + 1, npy] = -va[npx, npy - 2]
    ua_tl[-1, npy] = va_tl[0, npy - 2]; ua[-1, npy] = va[0, npy - 2]
    ua_tl[0, npy] = va_tl[0, npy - 1]; ua[0, npy] = va[0, npy - 1]
    if is_ == 1:
        for j in range(js - 1, je + 2):
            uc_tl[0, j] = c1 * utmp_tl[-2, j] + c2 * utmp_tl[-1, j] + c3 * utmp_tl[0, j]
            uc[0, j] = c1 * utmp[-2, j] + c2 * utmp[-1, j] + c3 * utmp[0, j]
            ut_tl[1, j], ut[1, j] = edge_interpolate4_tlm([ua[i, j] for i in range(-1, 3)], [ua_tl[i, j] for i in range(-1, 3)],
                                                          [dxa[i, j] for i in range(-1, 3)])
            if ut[1, j] > 0.:
                uc_tl[1, j] = sin_sg(0, j, 3) * ut_tl[1, j]
                uc[1, j] = ut[1, j] * sin_sg(0, j, 3)
            else:
                uc_tl[1, j] = sin_sg(1, j, 1) * ut_tl[1, j]
                uc[1, j] = ut[1, j] * sin_sg(1, j, 1)
            uc_tl[2, j] = c1 * utmp_tl[3, j] + c2 * utmp_tl[2, j] + c3 * utmp_tl[1, j]
            uc[2, j] = c1 * utmp[3, j] + c2 * utmp[2, j] + c3 * utmp[1, j]
            ut_tl[0, j] = rsin_u[0, j] * (uc_tl[0, j] - cosa_u[0, j] * v_tl[0, j])
            ut[0, j] = (uc[0, j] - v[0, j] * cosa_u[0, j]) * rsin_u[0, j]
            ut_tl[2, j] = rsin_u[2, j] * (uc_tl[2, j] - cosa_u[2, j] * v_tl[2, j])
            ut[2, j] = (uc[2, j] - v[2, j] * cosa_u[2, j]) * rsin_u[2, j]
    if ie + 1 == npx:
        for j in range(js - 1, je + 2):
            uc_tl[npx - 1, j] = c1 * utmp_tl[npx - 3, j] + c2 * utmp_tl[npx - 2, j] + c3 * utmp_tl[npx - 1, j]
            uc[npx - 1, j] = c1 * utmp[npx - 3, j] + c2 * utmp[npx - 2, j] + c3 * utmp[npx - 1, j]
            ut_tl[npx, j], ut[npx, j] = edge_interpolate4_tlm([ua[i, j] for i in range(npx - 2, npx + 2)],
                                                              [ua_tl[i, j] for i in range(npx - 2, npx + 2)],
                                                              [dxa[i, j] for i in range(npx - 2, npx + 2)])
            if ut[npx, j] > 0.:
                uc_tl[npx, j] = sin_sg(npx - 1, j, 3) * ut_tl[npx, j]
                uc[npx, j] = ut[npx, j] * sin_sg(npx - 1, j, 3)
            else:
                uc_tl[npx, j] = sin_sg(npx, j, 1) * ut_tl[npx, j]
                uc[npx, j] = ut[npx, j] * sin_sg(npx, j, 1)
            uc_tl[npx + 1, j] = c3 * utmp_tl[npx, j] + c2 * utmp_tl[npx + 1, j] + c1 * utmp_tl[npx + 2, j]
            uc[npx + 1, j] = c3 * utmp[npx, j] + c2 * utmp[npx + 1, j] + c1 * utmp[npx + 2, j]
            ut_tl[npx - 1, j] = rsin_u[npx - 1, j] * (uc_tl[npx - 1, j] - cosa_u[npx - 1, j] * v_tl[npx - 1, j])
            ut[npx - 1, j] = (uc[npx - 1, j] - v[npx - 1, j] * cosa_u[npx - 1, j]) * rsin_u[npx - 1, j]
            ut_tl[npx + 1, j] = rsin_u[npx + 1, j] * (uc_tl[npx + 1, j] - cosa_u[npx + 1, j] * v_tl[npx + 1, j])
            ut[npx + 1, j] = (uc[npx + 1, j] - v[npx + 1, j] * cosa_u[npx + 1, j]) * rsin_u[npx + 1, j]
    # Ydir
    for j in range(-2, 1):
        vtmp_tl[0, j] = -utmp_tl[1 - j, 0]
        vtmp[0, j] = -utmp[1 - j, 0]
    for j in range(0, 3):
        vtmp_tl[0, npy + j] = utmp_tl[j + 1, npy]
        vtmp[0, npy + j] = utmp[j + 1, npy]
    for j in range(-2, 1):
        vtmp_tl[npx, j] = utmp_tl[ie + j, 0]
        vtmp[npx, j] = utmp[ie + j, 0]
    for j in range(0, 3):
        vtmp_tl[npx, npy + j] = -utmp_tl[ie - j, npy]
        vtmp[npx, npy + j] = -utmp[ie - j, npy]
    va_tl[0, -1] = -ua_tl[2, 0]; va[0, -1] = -ua[2, 0]
    va_tl[0, 0] = -ua_tl[1, 0]; va[0, 0] = -ua[1, 0]
    va_tl[npx, 0] = ua_tl[npx - 1, 0]; va[npx, 0] = ua[npx - 1, 0]
    va_tl[npx, -1] = ua_tl[npx - 2, 0]; va[npx, -1] = ua[npx - 2, 0]
    va_tl[npx, npy] = -ua_tl[npx - 1, npy]; va[npx, npy] = -ua[npx - 1, npy]
    va_tl[npx, npy + 1] = -ua_tl[npx - 2, npy]; va[npx, npy + 1] = -ua[npx - 2, npy]
    va_tl[0, npy] = ua_tl[1, npy]; va[0, npy] = ua[1, npy]
    va_tl[0, npy + 1] = ua_tl[2, npy]; va[0, npy + 1] = ua[2, npy]
    for j in range(js - 1, je + 3):
        if j == 1 or j == npy:
            for i in range(is_ - 1, ie + 2):
                vt_tl[i, j], vt[i, j] = edge_interpolate4_tlm([va[i, jj] for jj in range(j - 2, j + 2)], [va_tl[i, jj] for jj in range(j - 2, j + 2)],
                                                              [dya[i, jj] for jj in range(j - 2, j + 2)])
                if vt[i, j] > 0.:
                    vc_tl[i, j] = sin_sg(i, j - 1, 4) * vt_tl[i, j]
                    vc[i, j] = vt[i, j] * sin_sg(i, j - 1, 4)
                else:
                    vc_tl[i, j] = sin_sg(i, j, 2) * vt_tl[i, j]
                    vc[i, j] = vt[i, j] * sin_sg(i, j, 2)
        elif j == 0 or j == npy - 1:
            for i in range(is_ - 1, ie + 2):
                vc_tl[i, j] = c1 * vtmp_tl[i, j - 2] + c2 * vtmp_tl[i, j - 1] + c3 * vtmp_tl[i, j]
                vc[i, j] = c1 * vtmp[i, j - 2] + c2 * vtmp[i, j - 1] + c3 * vtmp[i, j]
                vt_tl[i, j] = rsin_v[i, j] * (vc_tl[i, j] - cosa_v[i, j] * u_tl[i, j])
                vt[i, j] = (vc[i, j] - u[i, j] * cosa_v[i, j]) * rsin_v[i, j]
        elif j == 2 or j == npy + 1:
            for i in range(is_ - 1, ie + 2):
                vc_tl[i, j] = c1 * vtmp_tl[i, j + 1] + c2 * vtmp_tl[i, j] + c3 * vtmp_tl[i, j - 1]
                vc[i, j] = c1 * vtmp[i, j + 1] + c2 * vtmp[i, j] + c3 * vtmp[i, j - 1]
                vt_tl[i, j] = rsin_v[i, j] * (vc_tl[i, j] - cosa_v[i, j] * u_tl[i, j])
                vt[i, j] = (vc[i, j] - u[i, j] * cosa_v[i, j]) * rsin_v[i, j]
        else:
            for i in range(is_ - 1, ie + 2):
                vc_tl[i, j] = a2 * (vtmp_tl[i, j - 2] + vtmp_tl[i, j + 1]) + a1 * (vtmp_tl[i, j - 1] + vtmp_tl[i, j])
                vc[i, j] = a2 * (vtmp[i, j - 2] + vtmp[i, j + 1]) + a1 * (vtmp[i, j - 1] + vtmp[i, j])
                vt_tl[i, j] = rsin_v[i, j] * (vc_tl[i, j] - cosa_v[i, j] * u_tl[i, j])
                vt[i, j] = (vc[i, j] - u[i, j] * cosa_v[i, j]) * rsin_v[i, j]
    return dict(ua=ua, va=va, uc=uc, vc=vc, ut=ut, vt=vt, ua_tl=ua_tl, va_tl=va_tl, uc_tl=uc_tl, vc_tl=vc_tl, ut_tl=ut_tl, vt_tl=vt_tl)


def divergence_corner_tlm(u, u_tl, v, v_tl, ua, ua_tl, va, va_tl, gs, bd, npx, npy):
    """:3805-3966, the cubed-sphere branch.  gs: F arrays dxc, dyc, rarea_c and callables sin_sg, cos_sg.  Returns divg_d, divg_d_tl"""
    is_, ie, js, je = bd.is_, bd.ie, bd.js, bd.je
    sin_sg, cos_sg, dxc, dyc, rarea_c = gs["sin_sg"], gs["cos_sg"], gs["dxc"], gs["dyc"], gs["rarea_c"]
    divg_d = F((bd.isd, bd.ied + 1), (bd.jsd, bd.jed + 1)); divg_d_tl = F((bd.isd, bd.ied + 1), (bd.jsd, bd.jed + 1))
    uf = F((is_ - 2, ie + 2), (js - 1, je + 2)); uf_tl = F((is_ - 2, ie + 2), (js - 1, je + 2))
    vf = F((is_ - 1, ie + 2), (js - 2, je + 2)); vf_tl = F((is_ - 1, ie + 2), (js - 2, je + 2))
    is2 = is_ if 2 < is_ else 2
    ie1 = ie + 1 if npx - 1 > ie + 1 else npx - 1
    for j in range(js, je + 2):
        if j == 1 or j == npy:
            for i in range(is_ - 1, ie + 2):
                uf_tl[i, j] = dyc[i, j] * 0.5 * (sin_sg(i, j - 1, 4) + sin_sg(i, j, 2)) * u_tl[i, j]
                uf[i, j] = u[i, j] * dyc[i, j] * 0.5 * (sin_sg(i, j - 1, 4) + sin_sg(i, j, 2))
        else:
            for i in range(is_ - 1, ie + 2):
                uf_tl[i, j] = dyc[i, j] * 0.5 * (sin_sg(i, j - 1, 4) + sin_sg(i, j, 2)) * (
                    u_tl[i, j] - 0.25 * (cos_sg(i, j - 1, 4) + cos_sg(i, j, 2)) * (va_tl[i, j - 1] + va_tl[i, j]))
                uf[i, j] = ((u[i, j] - 0.25 * (va[i, j - 1] + va[i, j]) * (cos_sg(i, j - 1, 4) + cos_sg(i, j, 2))) * dyc[i, j] * 0.5
                            * (sin_sg(i, j - 1, 4) + sin_sg(i, j, 2)))
    for j in range(js - 1, je + 2):
        for i in range(is2, ie1 + 1):
            vf_tl[i, j] = dxc[i, j] * 0.5 * (sin_sg(i - 1, j, 3) + sin_sg(i, j, 1)) * (
                v_tl[i, j] - 0.25 * (cos_sg(i - 1, j, 3) + cos_sg(i, j, 1)) * (ua_tl[i - 1, j] + ua_tl[i, j]))
            vf[i, j] = ((v[i, j] - 0.25 * (ua[i - 1, j] + ua[i, j]) * (cos_sg(i - 1, j, 3) + cos_sg(i, j, 1))) * dxc[i, j] * 0.5
                        * (sin_sg(i - 1, j, 3) + sin_sg(i, j, 1)))
        if is_ == 1:
            vf_tl[1, j] = dxc[1, j] * 0.5 * (sin_sg(0, j, 3) + sin_sg(1, j, 1)) * v_tl[1, j]
            vf[1, j] = v[1, j] * dxc[1, j] * 0.5 * (sin_sg(0, j, 3) + sin_sg(1, j, 1))
        if ie + 1 == npx:
            vf_tl[npx, j] = dxc[npx, j] * 0.5 * (sin_sg(npx - 1, j, 3) + sin_sg(npx, j, 1)) * v_tl[npx, j]
            vf[npx, j] = v[npx, j] * dxc[npx, j] * 0.5 * (sin_sg(npx - 1, j, 3) + sin_sg(npx, j, 1))
    for j in range(js, je + 2):
        for i in range(is_, ie + 2):
            divg_d_tl[i, j] = vf_tl[i, j - 1] - vf_tl[i, j] + uf_tl[i - 1, j] - uf_tl[i, j]
            divg_d[i, j] = vf[i, j - 1] - vf[i, j] + (uf[i - 1, j] - uf[i, j])
    divg_d_tl[1, 1] = divg_d_tl[1, 1] - vf_tl[1, 0]
    divg_d[1, 1] = divg_d[1, 1] - vf[1, 0]
    divg_d_tl[npx, 1] = divg_d_tl[npx, 1] - vf_tl[npx, 0]
    divg_d[npx, 1] = divg_d[npx, 1] - vf[npx, 0]
    divg_d_tl[npx, npy] = divg_d_tl[npx, npy] + vf_tl[npx, npy]
    divg_d[npx, npy] = divg_d[npx, npy] + vf[npx, npy]
    divg_d_tl[1, npy] = divg_d_tl[1, npy] + vf_tl[1, npy]
    divg_d[1, npy] = divg_d[1, npy] + vf[1, npy]
    for j in range(js, je + 2):
        for i in range(is_, ie + 2):
            divg_d_tl[i, j] = rarea_c[i, j] * divg_d_tl[i, j]
            divg_d[i, j] = rarea_c[i, j] * divg_d[i, j]
    return divg_d, divg_d_tl
