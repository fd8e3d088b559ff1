"""A2B_ORD4_TLM, model_tlmadm/a2b_edge_tlm.F90:546-1158 (cubed-sphere branch: grid_type < 3, not nested, no `replace`) with
EXTRAP_CORNER_TLM (:1491-1504), transliterated statement by statement.  Fortran arrays are ref_tlm.F objects indexed (i, j)."""
import math
from . import F

r3 = 1. / 3.                                   # a2b_edge_tlm.F90:31
a1, a2 = 0.5625, -0.0625                       # :35-36
b1, b2 = 7. / 12., -1. / 12.                   # :40-41
c1, c2 = 2. / 3., -(1. / 6.)                   # local parameters of the routine


def great_circle_dist(q1, q2):
    """fv_grid_utils great_circle_dist without radius: the angle (haversine form)"""
    p1, p2 = q1, q2
    beta = math.asin(math.sqrt(math.sin((p1[1] - p2[1]) / 2.) ** 2 + math.cos(p1[1]) * math.cos(p2[1]) * math.sin((p1[0] - p2[0]) / 2.) ** 2)) * 2.
    return beta


def extrap_corner_tlm(p0, p1, p2, q1, q1_tl, q2, q2_tl):
    x1 = great_circle_dist(p1, p0)
    x2 = great_circle_dist(p2, p0)
    return q1_tl + x1 * (q1_tl - q2_tl) / (x2 - x1), q1 + x1 / (x2 - x1) * (q1 - q2)


def a2b_ord4_tlm(qin, qin_tl, grid, agrid, dxa, dya, edge_w, edge_e, edge_s, edge_n, npx, npy, is_, ie, js, je, ng,
                 sw_corner=True, se_corner=True, ne_corner=True, nw_corner=True):
    """qin, qin_tl, dxa, dya: F((is-ng, ie+ng), (js-ng, je+ng)); grid(i, j) / agrid(i, j) -> (lon, lat) callables; edge_*: callables of the
    Fortran index.  Returns qout, qout_tl as F over the same bounds (defined on is..ie+1, js..je+1)."""
    B = ((is_ - ng, ie + ng), (js - ng, je + ng))
    qout = F(*B); qout_tl = F(*B)
    qx = F((is_, ie + 1), (js - ng, je + ng)); qx_tl = F((is_, ie + 1), (js - ng, je + ng))
    qy = F((is_ - ng, ie + ng), (js, je + 1)); qy_tl = F((is_ - ng, ie + ng), (js, je + 1))
    qxx = F(*B); qxx_tl = F(*B); qyy = F(*B); qyy_tl = F(*B)
    q1 = F((is_ - 1, ie + 1)); q1_tl = F((is_ - 1, ie + 1)); q2 = F((js - 1, je + 1)); q2_tl = F((js - 1, je + 1))
    is1 = is_ - 1 if 1 < is_ - 1 else 1
    js1 = js - 1 if 1 < js - 1 else 1
    is2 = is_ if 2 < is_ else 2
    js2 = js if 2 < js else 2
    ie1 = ie + 1 if npx - 1 > ie + 1 else npx - 1
    je1 = je + 1 if npy - 1 > je + 1 else npy - 1
    # Corners: 3-way extrapolation
    if sw_corner:
        p0 = grid(1, 1)
        result1_tl, result1 = extrap_corner_tlm(p0, agrid(1, 1), agrid(2, 2), qin[1, 1], qin_tl[1, 1], qin[2, 2], qin_tl[2, 2])
        result2_tl, result2 = extrap_corner_tlm(p0, agrid(0, 1), agrid(-1, 2), qin[0, 1], qin_tl[0, 1], qin[-1, 2], qin_tl[-1, 2])
        result3_tl, result3 = extrap_corner_tlm(p0, agrid(1, 0), agrid(2, -1), qin[1, 0], qin_tl[1, 0], qin[2, -1], qin_tl[2, -1])
        qout_tl[1, 1] = r3 * (result1_tl + result2_tl + result3_tl)
        qout[1, 1] = (result1 + result2 + result3) * r3
    if se_corner:
        p0 = grid(npx, 1)
        result1_tl, result1 = extrap_corner_tlm(p0, agrid(npx - 1, 1), agrid(npx - 2, 2), qin[npx - 1, 1], qin_tl[npx - 1, 1], qin[npx - 2, 2], qin_tl[npx - 2, 2])
        result2_tl, result2 = extrap_corner_tlm(p0, agrid(npx - 1, 0), agrid(npx - 2, -1), qin[npx - 1, 0], qin_tl[npx - 1, 0], qin[npx - 2, -1], qin_tl[npx - 2, -1])
        result3_tl, result3 = extrap_corner_tlm(p0, agrid(npx, 1), agrid(npx + 1, 2), qin[npx, 1], qin_tl[npx, 1], qin[npx + 1, 2], qin_tl[npx + 1, 2])
        qout_tl[npx, 1] = r3 * (result1_tl + result2_tl + result3_tl)
        qout[npx, 1] = (result1 + result2 + result3) * r3
    if ne_corner:
        p0 = grid(npx, npy)
        result1_tl, result1 = extrap_corner_tlm(p0, agrid(npx - 1, npy - 1), agrid(npx - 2, npy - 2), qin[npx - 1, npy - 1], qin_tl[npx - 1, npy - 1],
                                                qin[npx - 2, npy - 2], qin_tl[npx - 2, npy - 2])
        result2_tl, result2 = extrap_corner_tlm(p0, agrid(npx, npy - 1), agrid(npx + 1, npy - 2), qin[npx, npy - 1], qin_tl[npx, npy - 1],
                                                qin[npx + 1, npy - 2], qin_tl[npx + 1, npy - 2])
        result3_tl, result3 = extrap_corner_tlm(p0, agrid(npx - 1, npy), agrid(npx - 2, npy + 1), qin[npx - 1, npy], qin_tl[npx - 1, npy],
                                                qin[npx - 2, npy + 1], qin_tl[npx - 2, npy + 1])
        qout_tl[npx, npy] = r3 * (result1_tl + result2_tl + result3_tl)
        qout[npx, npy] = (result1 + result2 + result3) * r3
    if nw_corner:
        p0 = grid(1, npy)
        result1_tl, result1 = extrap_corner_tlm(p0, agrid(1, npy - 1), agrid(2, npy - 2), qin[1, npy - 1], qin_tl[1, npy - 1], qin[2, npy - 2], qin_tl[2, npy - 2])
        result2_tl, result2 = extrap_corner_tlm(p0, agrid(0, npy - 1), agrid(-1, npy - 2), qin[0, npy - 1], qin_tl[0, npy - 1], qin[-1, npy - 2], qin_tl[-1, npy - 2])
        result3_tl, result3 = extrap_corner_tlm(p0, agrid(1, npy), agrid(2, npy + 1), qin[1, npy], qin_tl[1, npy], qin[2, npy + 1], qin_tl[2, npy + 1])
        qout_tl[1, npy] = r3 * (result1_tl + result2_tl + result3_tl)
        qout[1, npy] = (result1 + result2 + result3) * r3
    max1 = js - 2 if 1 < js - 2 else 1
    min1 = je + 2 if npy - 1 > je + 2 else npy - 1
    qx_tl.fill(0.0)
    # X-Interior:
    for j in range(max1, min1 + 1):
        max2 = is_ if 3 < is_ else 3
        min2 = ie + 1 if npx - 2 > ie + 1 else npx - 2
        for i in range(max2, min2 + 1):
            qx_tl[i, j] = b2 * (qin_tl[i - 2, j] + qin_tl[i + 1, j]) + b1 * (qin_tl[i - 1, j] + qin_tl[i, j])
            qx[i, j] = b2 * (qin[i - 2, j] + qin[i + 1, j]) + b1 * (qin[i - 1, j] + qin[i, j])
    # *** West Edges:
    if is_ == 1:
        q2_tl.fill(0.0)
        for j in range(js1, je1 + 1):
            q2_tl[j] = (dxa[1, j] * qin_tl[0, j] + dxa[0, j] * qin_tl[1, j]) / (dxa[0, j] + dxa[1, j])
            q2[j] = (qin[0, j] * dxa[1, j] + qin[1, j] * dxa[0, j]) / (dxa[0, j] + dxa[1, j])
        for j in range(js2, je1 + 1):
            qout_tl[1, j] = edge_w(j) * q2_tl[j - 1] + (1. - edge_w(j)) * q2_tl[j]
            qout[1, j] = edge_w(j) * q2[j - 1] + (1. - edge_w(j)) * q2[j]
        max3 = js - 2 if 1 < js - 2 else 1
        min3 = je + 2 if npy - 1 > je + 2 else npy - 1
        for j in range(max3, min3 + 1):
            g_in = dxa[2, j] / dxa[1, j]
            g_ou = dxa[-1, j] / dxa[0, j]
            qx_tl[1, j] = 0.5 * (((2. + g_in) * qin_tl[1, j] - qin_tl[2, j]) / (1. + g_in) + ((2. + g_ou) * qin_tl[0, j] - qin_tl[-1, j]) / (1. + g_ou))
            qx[1, j] = 0.5 * (((2. + g_in) * qin[1, j] - qin[2, j]) / (1. + g_in) + ((2. + g_ou) * qin[0, j] - qin[-1, j]) / (1. + g_ou))
            qx_tl[2, j] = (3. * (g_in * qin_tl[1, j] + qin_tl[2, j]) - g_in * qx_tl[1, j] - qx_tl[3, j]) / (2. + 2. * g_in)
            qx[2, j] = (3. * (g_in * qin[1, j] + qin[2, j]) - (g_in * qx[1, j] + qx[3, j])) / (2. + 2. * g_in)
    else:
        q2_tl.fill(0.0)
    # East Edges:
    if ie + 1 == npx:
        for j in range(js1, je1 + 1):
            q2_tl[j] = (dxa[npx, j] * qin_tl[npx - 1, j] + dxa[npx - 1, j] * qin_tl[npx, j]) / (dxa[npx - 1, j] + dxa[npx, j])
            q2[j] = (qin[npx - 1, j] * dxa[npx, j] + qin[npx, j] * dxa[npx - 1, j]) / (dxa[npx - 1, j] + dxa[npx, j])
        for j in range(js2, je1 + 1):
            qout_tl[npx, j] = edge_e(j) * q2_tl[j - 1] + (1. - edge_e(j)) * q2_tl[j]
            qout[npx, j] = edge_e(j) * q2[j - 1] + (1. - edge_e(j)) * q2[j]
        max4 = js - 2 if 1 < js - 2 else 1
        min4 = je + 2 if npy - 1 > je + 2 else npy - 1
        for j in range(max4, min4 + 1):
            g_in = dxa[npx - 2, j] / dxa[npx - 1, j]
            g_ou = dxa[npx + 1, j] / dxa[npx, j]
            qx_tl[npx, j] = 0.5 * (((2. + g_in) * qin_tl[npx - 1, j] - qin_tl[npx - 2, j]) / (1. + g_in)
                                   + ((2. + g_ou) * qin_tl[npx, j] - qin_tl[npx + 1, j]) / (1. + g_ou))
            qx[npx, j] = 0.5 * (((2. + g_in) * qin[npx - 1, j] - qin[npx - 2, j]) / (1. + g_in) + ((2. + g_ou) * qin[npx, j] - qin[npx + 1, j]) / (1. + g_ou))
            qx_tl[npx - 1, j] = (3. * (qin_tl[npx - 2, j] + g_in * qin_tl[npx - 1, j]) - g_in * qx_tl[npx, j] - qx_tl[npx - 2, j]) / (2. + 2. * g_in)
            qx[npx - 1, j] = (3. * (qin[npx - 2, j] + g_in * qin[npx - 1, j]) - (g_in * qx[npx, j] + qx[npx - 2, j])) / (2. + 2. * g_in)
    # Y-Interior:
    max5 = js if 3 < js else 3
    min5 = je + 1 if npy - 2 > je + 1 else npy - 2
    qy_tl.fill(0.0)
    for j in range(max5, min5 + 1):
        max6 = is_ - 2 if 1 < is_ - 2 else 1
        min6 = ie + 2 if npx - 1 > ie + 2 else npx - 1
        for i in range(max6, min6 + 1):
            qy_tl[i, j] = b2 * (qin_tl[i, j - 2] + qin_tl[i, j + 1]) + b1 * (qin_tl[i, j - 1] + qin_tl[i, j])
            qy[i, j] = b2 * (qin[i, j - 2] + qin[i, j + 1]) + b1 * (qin[i, j - 1] + qin[i, j])
    # South Edges:
    if js == 1:
        q1_tl.fill(0.0)
        for i in range(is1, ie1 + 1):
            q1_tl[i] = (dya[i, 1] * qin_tl[i, 0] + dya[i, 0] * qin_tl[i, 1]) / (dya[i, 0] + dya[i, 1])
            q1[i] = (qin[i, 0] * dya[i, 1] + qin[i, 1] * dya[i, 0]) / (dya[i, 0] + dya[i, 1])
        for i in range(is2, ie1 + 1):
            qout_tl[i, 1] = edge_s(i) * q1_tl[i - 1] + (1. - edge_s(i)) * q1_tl[i]
            qout[i, 1] = edge_s(i) * q1[i - 1] + (1. - edge_s(i)) * q1[i]
        max7 = is_ - 2 if 1 < is_ - 2 else 1
        min7 = ie + 2 if npx - 1 > ie + 2 else npx - 1
        for i in range(max7, min7 + 1):
            g_in = dya[i, 2] / dya[i, 1]
            g_ou = dya[i, -1] / dya[i, 0]
            qy_tl[i, 1] = 0.5 * (((2. + g_in) * qin_tl[i, 1] - qin_tl[i, 2]) / (1. + g_in) + ((2. + g_ou) * qin_tl[i, 0] - qin_tl[i, -1]) / (1. + g_ou))
            qy[i, 1] = 0.5 * (((2. + g_in) * qin[i, 1] - qin[i, 2]) / (1. + g_in) + ((2. + g_ou) * qin[i, 0] - qin[i, -1]) / (1. + g_ou))
            qy_tl[i, 2] = (3. * (g_in * qin_tl[i, 1] + qin_tl[i, 2]) - g_in * qy_tl[i, 1] - qy_tl[i, 3]) / (2. + 2. * g_in)
            qy[i, 2] = (3. * (g_in * qin[i, 1] + qin[i, 2]) - (g_in * qy[i, 1] + qy[i, 3])) / (2. + 2. * g_in)
    else:
        q1_tl.fill(0.0)
    # North Edges:
    if je + 1 == npy:
        for i in range(is1, ie1 + 1):
            q1_tl[i] = (dya[i, npy] * qin_tl[i, npy - 1] + dya[i, npy - 1] * qin_tl[i, npy]) / (dya[i, npy - 1] + dya[i, npy])
            q1[i] = (qin[i, npy - 1] * dya[i, npy] + qin[i, npy] * dya[i, npy - 1]) / (dya[i, npy - 1] + dya[i, npy])
        for i in range(is2, ie1 + 1):
            qout_tl[i, npy] = edge_n(i) * q1_tl[i - 1] + (1. - edge_n(i)) * q1_tl[i]
            qout[i, npy] = edge_n(i) * q1[i - 1] + (1. - edge_n(i)) * q1[i]
        max8 = is_ - 2 if 1 < is_ - 2 else 1
        min8 = ie + 2 if npx - 1 > ie + 2 else npx - 1
        for i in range(max8, min8 + 1):
            g_in = dya[i, npy - 2] / dya[i, npy - 1]
            g_ou = dya[i, npy + 1] / dya[i, npy]
            qy_tl[i, npy] = 0.5 * (((2. + g_in) * qin_tl[i, npy - 1] - qin_tl[i, npy - 2]) / (1. + g_in)
                                   + ((2. + g_ou) * qin_tl[i, npy] - qin_tl[i, npy + 1]) / (1. + g_ou))
            qy[i, npy] = 0.5 * (((2. + g_in) * qin[i, npy - 1] - qin[i, npy - 2]) / (1. + g_in) + ((2. + g_ou) * qin[i, npy] - qin[i, npy + 1]) / (1. + g_ou))
            qy_tl[i, npy - 1] = (3. * (qin_tl[i, npy - 2] + g_in * qin_tl[i, npy - 1]) - g_in * qy_tl[i, npy] - qy_tl[i, npy - 2]) / (2. + 2. * g_in)
            qy[i, npy - 1] = (3. * (qin[i, npy - 2] + g_in * qin[i, npy - 1]) - (g_in * qy[i, npy] + qy[i, npy - 2])) / (2. + 2. * g_in)
    # --------------------------------------
    max9 = js if 3 < js else 3
    min9 = je + 1 if npy - 2 > je + 1 else npy - 2
    qxx_tl.fill(0.0)
    for j in range(max9, min9 + 1):
        max10 = is_ if 2 < is_ else 2
        min10 = ie + 1 if npx - 1 > ie + 1 else npx - 1
        for i in range(max10, min10 + 1):
            qxx_tl[i, j] = a2 * (qx_tl[i, j - 2] + qx_tl[i, j + 1]) + a1 * (qx_tl[i, j - 1] + qx_tl[i, j])
            qxx[i, j] = a2 * (qx[i, j - 2] + qx[i, j + 1]) + a1 * (qx[i, j - 1] + qx[i, j])
    if js == 1:
        max11 = is_ if 2 < is_ else 2
        min11 = ie + 1 if npx - 1 > ie + 1 else npx - 1
        for i in range(max11, min11 + 1):
            qxx_tl[i, 2] = c1 * (qx_tl[i, 1] + qx_tl[i, 2]) + c2 * (qout_tl[i, 1] + qxx_tl[i, 3])
            qxx[i, 2] = c1 * (qx[i, 1] + qx[i, 2]) + c2 * (qout[i, 1] + qxx[i, 3])
    if je + 1 == npy:
        max12 = is_ if 2 < is_ else 2
        min12 = ie + 1 if npx - 1 > ie + 1 else npx - 1
        for i in range(max12, min12 + 1):
            qxx_tl[i, npy - 1] = c1 * (qx_tl[i, npy - 2] + qx_tl[i, npy - 1]) + c2 * (qout_tl[i, npy] + qxx_tl[i, npy - 2])
            qxx[i, npy - 1] = c1 * (qx[i, npy - 2] + qx[i, npy - 1]) + c2 * (qout[i, npy] + qxx[i, npy - 2])
    max13 = js if 2 < js else 2
    min13 = je + 1 if npy - 1 > je + 1 else npy - 1
    qyy_tl.fill(0.0)
    for j in range(max13, min13 + 1):
        max14 = is_ if 3 < is_ else 3
        min14 = ie + 1 if npx - 2 > ie + 1 else npx - 2
        for i in range(max14, min14 + 1):
            qyy_tl[i, j] = a2 * (qy_tl[i - 2, j] + qy_tl[i + 1, j]) + a1 * (qy_tl[i - 1, j] + qy_tl[i, j])
            qyy[i, j] = a2 * (qy[i - 2, j] + qy[i + 1, j]) + a1 * (qy[i - 1, j] + qy[i, j])
        if is_ == 1:
            qyy_tl[2, j] = c1 * (qy_tl[1, j] + qy_tl[2, j]) + c2 * (qout_tl[1, j] + qyy_tl[3, j])
            qyy[2, j] = c1 * (qy[1, j] + qy[2, j]) + c2 * (qout[1, j] + qyy[3, j])
        if ie + 1 == npx:
            qyy_tl[npx - 1, j] = c1 * (qy_tl[npx - 2, j] + qy_tl[npx - 1, j]) + c2 * (qout_tl[npx, j] + qyy_tl[npx - 2, j])
            qyy[npx - 1, j] = c1 * (qy[npx - 2, j] + qy[npx - 1, j]) + c2 * (qout[npx, j] + qyy[npx - 2, j])
        max15 = is_ if 2 < is_ else 2
        min15 = ie + 1 if npx - 1 > ie + 1 else npx - 1
        for i in range(max15, min15 + 1):
            # averaging
            qout_tl[i, j] = 0.5 * (qxx_tl[i, j] + qyy_tl[i, j])
            qout[i, j] = 0.5 * (qxx[i, j] + qyy[i, j])
    return qout, qout_tl
