"""XTP_U_TLM (model_tlmadm/sw_core_tlm.F90:7272-7485) and YTP_V_TLM (:7490-7757), transliterated statement by statement for a whole cube tile
(is = js = 1, ie + 1 = npx, je + 1 = npy, not nested, grid_type 0): the linear orders the tangent model has (1, 2, 333)."""
from . import F
from .xppm_tlm import p1, p2, c1, c2, c3


def xtp_u_tlm(is_, ie, js, je, isd, ied, jsd, jed, c, c_tl, u, u_tl, iord, dx, rdx, npx, npy):
    flux = F((is_, ie + 1), (js, je + 1)); flux_tl = F((is_, ie + 1), (js, je + 1))
    bl = F((is_ - 1, ie + 1)); br = F((is_ - 1, ie + 1)); b0 = F((is_ - 1, ie + 1))
    bl_tl = F((is_ - 1, ie + 1)); br_tl = F((is_ - 1, ie + 1)); b0_tl = F((is_ - 1, ie + 1))
    al = F((is_ - 1, ie + 2)); al_tl = F((is_ - 1, ie + 2))
    is3 = is_ - 1 if 3 < is_ - 1 else 3
    ie3 = ie + 1 if npx - 3 > ie + 1 else npx - 3
    if iord == 1:
        for j in range(js, je + 2):
            for i in range(is_, ie + 2):
                if c[i, j] > 0.:
                    flux_tl[i, j] = u_tl[i - 1, j]
                    flux[i, j] = u[i - 1, j]
                else:
                    flux_tl[i, j] = u_tl[i, j]
                    flux[i, j] = u[i, j]
    elif iord == 333:
        for j in range(js, je + 2):
            for i in range(is_, ie + 2):
                if c[i, j] > 0.:
                    flux_tl[i, j] = ((2.0 * u_tl[i, j] + 5.0 * u_tl[i - 1, j] - u_tl[i - 2, j]) / 6.0
                                     - 0.5 * rdx[i - 1, j] * (c_tl[i, j] * (u[i, j] - u[i - 1, j]) + c[i, j] * (u_tl[i, j] - u_tl[i - 1, j]))
                                     + rdx[i - 1, j] ** 2 * (c_tl[i, j] * c[i, j] + c[i, j] * c_tl[i, j]) * (u[i, j] - 2.0 * u[i - 1, j] + u[i - 2, j]) / 6.0
                                     + c[i, j] ** 2 * rdx[i - 1, j] ** 2 * (u_tl[i, j] - 2.0 * u_tl[i - 1, j] + u_tl[i - 2, j]) / 6.0)
                    flux[i, j] = ((2.0 * u[i, j] + 5.0 * u[i - 1, j] - u[i - 2, j]) / 6.0 - 0.5 * c[i, j] * rdx[i - 1, j] * (u[i, j] - u[i - 1, j])
                                  + c[i, j] * rdx[i - 1, j] * c[i, j] * rdx[i - 1, j] / 6.0 * (u[i, j] - 2.0 * u[i - 1, j] + u[i - 2, j]))
                else:
                    flux_tl[i, j] = ((2.0 * u_tl[i - 1, j] + 5.0 * u_tl[i, j] - u_tl[i + 1, j]) / 6.0
                                     - 0.5 * rdx[i, j] * (c_tl[i, j] * (u[i, j] - u[i - 1, j]) + c[i, j] * (u_tl[i, j] - u_tl[i - 1, j]))
                                     + rdx[i, j] ** 2 * (c_tl[i, j] * c[i, j] + c[i, j] * c_tl[i, j]) * (u[i + 1, j] - 2.0 * u[i, j] + u[i - 1, j]) / 6.0
                                     + c[i, j] ** 2 * rdx[i, j] ** 2 * (u_tl[i + 1, j] - 2.0 * u_tl[i, j] + u_tl[i - 1, j]) / 6.0)
                    flux[i, j] = ((2.0 * u[i - 1, j] + 5.0 * u[i, j] - u[i + 1, j]) / 6.0 - 0.5 * c[i, j] * rdx[i, j] * (u[i, j] - u[i - 1, j])
                                  + c[i, j] * rdx[i, j] * c[i, j] * rdx[i, j] / 6.0 * (u[i + 1, j] - 2.0 * u[i, j] + u[i - 1, j]))
    elif iord < 8:
        for j in range(js, je + 2):
            for i in range(is3, ie3 + 2):
                al_tl[i] = p1 * (u_tl[i - 1, j] + u_tl[i, j]) + p2 * (u_tl[i - 2, j] + u_tl[i + 1, j])
                al[i] = p1 * (u[i - 1, j] + u[i, j]) + p2 * (u[i - 2, j] + u[i + 1, j])
            for i in range(is3, ie3 + 1):
                bl_tl[i] = al_tl[i] - u_tl[i, j]
                bl[i] = al[i] - u[i, j]
                br_tl[i] = al_tl[i + 1] - u_tl[i, j]
                br[i] = al[i + 1] - u[i, j]
            if is_ == 1:
                xt_tl = c3 * u_tl[1, j] + c2 * u_tl[2, j] + c1 * u_tl[3, j]
                xt = c3 * u[1, j] + c2 * u[2, j] + c1 * u[3, j]
                br_tl[1] = xt_tl - u_tl[1, j]
                br[1] = xt - u[1, j]
                bl_tl[2] = xt_tl - u_tl[2, j]
                bl[2] = xt - u[2, j]
                br_tl[2] = al_tl[3] - u_tl[2, j]
                br[2] = al[3] - u[2, j]
                if j == 1 or j == npy:
                    bl_tl[0] = 0.0; bl[0] = 0.
                    br_tl[0] = 0.0; br[0] = 0.
                    bl_tl[1] = 0.0; bl[1] = 0.
                    br_tl[1] = 0.0; br[1] = 0.
                else:
                    bl_tl[0] = c1 * u_tl[-2, j] + c2 * u_tl[-1, j] + c3 * u_tl[0, j] - u_tl[0, j]
                    bl[0] = c1 * u[-2, j] + c2 * u[-1, j] + c3 * u[0, j] - u[0, j]
                    xt_tl = 0.5 * (((2. * dx[0, j] + dx[-1, j]) * u_tl[0, j] - dx[0, j] * u_tl[-1, j]) / (dx[0, j] + dx[-1, j])
                                   + ((2. * dx[1, j] + dx[2, j]) * u_tl[1, j] - dx[1, j] * u_tl[2, j]) / (dx[1, j] + dx[2, j]))
                    xt = 0.5 * (((2. * dx[0, j] + dx[-1, j]) * u[0, j] - dx[0, j] * u[-1, j]) / (dx[0, j] + dx[-1, j])
                                + ((2. * dx[1, j] + dx[2, j]) * u[1, j] - dx[1, j] * u[2, j]) / (dx[1, j] + dx[2, j]))
                    br_tl[0] = xt_tl - u_tl[0, j]
                    br[0] = xt - u[0, j]
                    bl_tl[1] = xt_tl - u_tl[1, j]
                    bl[1] = xt - u[1, j]
            if ie + 1 == npx:
                bl_tl[npx - 2] = al_tl[npx - 2] - u_tl[npx - 2, j]
                bl[npx - 2] = al[npx - 2] - u[npx - 2, j]
                xt_tl = c1 * u_tl[npx - 3, j] + c2 * u_tl[npx - 2, j] + c3 * u_tl[npx - 1, j]
                xt = c1 * u[npx - 3, j] + c2 * u[npx - 2, j] + c3 * u[npx - 1, j]
                br_tl[npx - 2] = xt_tl - u_tl[npx - 2, j]
                br[npx - 2] = xt - u[npx - 2, j]
                bl_tl[npx - 1] = xt_tl - u_tl[npx - 1, j]
                bl[npx - 1] = xt - u[npx - 1, j]
                if j == 1 or j == npy:
                    bl_tl[npx - 1] = 0.0; bl[npx - 1] = 0.
                    br_tl[npx - 1] = 0.0; br[npx - 1] = 0.
                    bl_tl[npx] = 0.0; bl[npx] = 0.
                    br_tl[npx] = 0.0; br[npx] = 0.
                else:
                    xt_tl = 0.5 * (((2. * dx[npx - 1, j] + dx[npx - 2, j]) * u_tl[npx - 1, j] - dx[npx - 1, j] * u_tl[npx - 2, j])
                                   / (dx[npx - 1, j] + dx[npx - 2, j])
                                   + ((2. * dx[npx, j] + dx[npx + 1, j]) * u_tl[npx, j] - dx[npx, j] * u_tl[npx + 1, j])
                                   / (dx[npx, j] + dx[npx + 1, j]))
                    xt = 0.5 * (((2. * dx[npx - 1, j] + dx[npx - 2, j]) * u[npx - 1, j] - dx[npx - 1, j] * u[npx - 2, j])
                                / (dx[npx - 1, j] + dx[npx - 2, j])
                                + ((2. * dx[npx, j] + dx[npx + 1, j]) * u[npx, j] - dx[npx, j] * u[npx + 1, j])
                                / (dx[npx, j] + dx[npx + 1, j]))
                    br_tl[npx - 1] = xt_tl - u_tl[npx - 1, j]
                    br[npx - 1] = xt - u[npx - 1, j]
                    bl_tl[npx] = xt_tl - u_tl[npx, j]
                    bl[npx] = xt - u[npx, j]
                    br_tl[npx] = c3 * u_tl[npx, j] + c2 * u_tl[npx + 1, j] + c1 * u_tl[npx + 2, j] - u_tl[npx, j]
                    br[npx] = c3 * u[npx, j] + c2 * u[npx + 1, j] + c1 * u[npx + 2, j] - u[npx, j]
            for i in range(is_ - 1, ie + 2):
                b0_tl[i] = bl_tl[i] + br_tl[i]
                b0[i] = bl[i] + br[i]
            if iord == 2:
                for i in range(is_, ie + 2):
                    if c[i, j] > 0.:
                        cfl_tl = rdx[i - 1, j] * c_tl[i, j]
                        cfl = c[i, j] * rdx[i - 1, j]
                        flux_tl[i, j] = (u_tl[i - 1, j] + (1. - cfl) * (br_tl[i - 1] - cfl_tl * b0[i - 1] - cfl * b0_tl[i - 1])
                                         - cfl_tl * (br[i - 1] - cfl * b0[i - 1]))
                        flux[i, j] = u[i - 1, j] + (1. - cfl) * (br[i - 1] - cfl * b0[i - 1])
                    else:
                        cfl_tl = rdx[i, j] * c_tl[i, j]
                        cfl = c[i, j] * rdx[i, j]
                        flux_tl[i, j] = (u_tl[i, j] + cfl_tl * (bl[i] + cfl * b0[i])
                                         + (1. + cfl) * (bl_tl[i] + cfl_tl * b0[i] + cfl * b0_tl[i]))
                        flux[i, j] = u[i, j] + (1. + cfl) * (bl[i] + cfl * b0[i])
    return flux, flux_tl


def ytp_v_tlm(is_, ie, js, je, isd, ied, jsd, jed, c, c_tl, v, v_tl, jord, dy, rdy, npx, npy):
    flux = F((is_, ie + 1), (js, je + 1)); flux_tl = F((is_, ie + 1), (js, je + 1))
    al = F((is_, ie + 1), (js - 1, je + 2)); al_tl = F((is_, ie + 1), (js - 1, je + 2))
    bl = F((is_, ie + 1), (js - 1, je + 1)); br = F((is_, ie + 1), (js - 1, je + 1)); b0 = F((is_, ie + 1), (js - 1, je + 1))
    bl_tl = F((is_, ie + 1), (js - 1, je + 1)); br_tl = F((is_, ie + 1), (js - 1, je + 1)); b0_tl = F((is_, ie + 1), (js - 1, je + 1))
    js3 = js - 1 if 3 < js - 1 else 3
    je3 = je + 1 if npy - 3 > je + 1 else npy - 3
    if jord == 1:
        for j in range(js, je + 2):
            for i in range(is_, ie + 2):
                if c[i, j] > 0.:
                    flux_tl[i, j] = v_tl[i, j - 1]
                    flux[i, j] = v[i, j - 1]
                else:
                    flux_tl[i, j] = v_tl[i, j]
                    flux[i, j] = v[i, j]
    elif jord == 333:
        for j in range(js, je + 2):
            for i in range(is_, ie + 2):
                if c[i, j] > 0.:
                    flux_tl[i, j] = ((2.0 * v_tl[i, j] + 5.0 * v_tl[i, j - 1] - v_tl[i, j - 2]) / 6.0
                                     - 0.5 * rdy[i, j - 1] * (c_tl[i, j] * (v[i, j] - v[i, j - 1]) + c[i, j] * (v_tl[i, j] - v_tl[i, j - 1]))
                                     + rdy[i, j - 1] ** 2 * (c_tl[i, j] * c[i, j] + c[i, j] * c_tl[i, j]) * (v[i, j] - 2.0 * v[i, j - 1] + v[i, j - 2]) / 6.0
                                     + c[i, j] ** 2 * rdy[i, j - 1] ** 2 * (v_tl[i, j] - 2.0 * v_tl[i, j - 1] + v_tl[i, j - 2]) / 6.0)
                    flux[i, j] = ((2.0 * v[i, j] + 5.0 * v[i, j - 1] - v[i, j - 2]) / 6.0 - 0.5 * c[i, j] * rdy[i, j - 1] * (v[i, j] - v[i, j - 1])
                                  + c[i, j] * rdy[i, j - 1] * c[i, j] * rdy[i, j - 1] / 6.0 * (v[i, j] - 2.0 * v[i, j - 1] + v[i, j - 2]))
                else:
                    flux_tl[i, j] = ((2.0 * v_tl[i, j - 1] + 5.0 * v_tl[i, j] - v_tl[i, j + 1]) / 6.0
                                     - 0.5 * rdy[i, j] * (c_tl[i, j] * (v[i, j] - v[i, j - 1]) + c[i, j] * (v_tl[i, j] - v_tl[i, j - 1]))
                                     + rdy[i, j] ** 2 * (c_tl[i, j] * c[i, j] + c[i, j] * c_tl[i, j]) * (v[i, j + 1] - 2.0 * v[i, j] + v[i, j - 1]) / 6.0
                                     + c[i, j] ** 2 * rdy[i, j] ** 2 * (v_tl[i, j + 1] - 2.0 * v_tl[i, j] + v_tl[i, j - 1]) / 6.0)
                    flux[i, j] = ((2.0 * v[i, j - 1] + 5.0 * v[i, j] - v[i, j + 1]) / 6.0 - 0.5 * c[i, j] * rdy[i, j] * (v[i, j] - v[i, j - 1])
                                  + c[i, j] * rdy[i, j] * c[i, j] * rdy[i, j] / 6.0 * (v[i, j + 1] - 2.0 * v[i, j] + v[i, j - 1]))
    elif jord < 8:
        for j in range(js3, je3 + 2):
            for i in range(is_, ie + 2):
                al_tl[i, j] = p1 * (v_tl[i, j - 1] + v_tl[i, j]) + p2 * (v_tl[i, j - 2] + v_tl[i, j + 1])
                al[i, j] = p1 * (v[i, j - 1] + v[i, j]) + p2 * (v[i, j - 2] + v[i, j + 1])
        for j in range(js3, je3 + 1):
            for i in range(is_, ie + 2):
                bl_tl[i, j] = al_tl[i, j] - v_tl[i, j]
                bl[i, j] = al[i, j] - v[i, j]
                br_tl[i, j] = al_tl[i, j + 1] - v_tl[i, j]
                br[i, j] = al[i, j + 1] - v[i, j]
        if js == 1:
            for i in range(is_, ie + 2):
                bl_tl[i, 0] = c1 * v_tl[i, -2] + c2 * v_tl[i, -1] + c3 * v_tl[i, 0] - v_tl[i, 0]
                bl[i, 0] = c1 * v[i, -2] + c2 * v[i, -1] + c3 * v[i, 0] - v[i, 0]
                xt_tl = 0.5 * (((2. * dy[i, 0] + dy[i, -1]) * v_tl[i, 0] - dy[i, 0] * v_tl[i, -1]) / (dy[i, 0] + dy[i, -1])
                               + ((2. * dy[i, 1] + dy[i, 2]) * v_tl[i, 1] - dy[i, 1] * v_tl[i, 2]) / (dy[i, 1] + dy[i, 2]))
                xt = 0.5 * (((2. * dy[i, 0] + dy[i, -1]) * v[i, 0] - dy[i, 0] * v[i, -1]) / (dy[i, 0] + dy[i, -1])
                            + ((2. * dy[i, 1] + dy[i, 2]) * v[i, 1] - dy[i, 1] * v[i, 2]) / (dy[i, 1] + dy[i, 2]))
                br_tl[i, 0] = xt_tl - v_tl[i, 0]
                br[i, 0] = xt - v[i, 0]
                bl_tl[i, 1] = xt_tl - v_tl[i, 1]
                bl[i, 1] = xt - v[i, 1]
                xt_tl = c3 * v_tl[i, 1] + c2 * v_tl[i, 2] + c1 * v_tl[i, 3]
                xt = c3 * v[i, 1] + c2 * v[i, 2] + c1 * v[i, 3]
                br_tl[i, 1] = xt_tl - v_tl[i, 1]
                br[i, 1] = xt - v[i, 1]
                bl_tl[i, 2] = xt_tl - v_tl[i, 2]
                bl[i, 2] = xt - v[i, 2]
                br_tl[i, 2] = al_tl[i, 3] - v_tl[i, 2]
                br[i, 2] = al[i, 3] - v[i, 2]
            if is_ == 1:
                for jj in (0, 1):
                    bl_tl[1, jj] = 0.0; bl[1, jj] = 0.
                    br_tl[1, jj] = 0.0; br[1, jj] = 0.
            if ie + 1 == npx:
                for jj in (0, 1):
                    bl_tl[npx, jj] = 0.0; bl[npx, jj] = 0.
                    br_tl[npx, jj] = 0.0; br[npx, jj] = 0.
        if je + 1 == npy:
            for i in range(is_, ie + 2):
                bl_tl[i, npy - 2] = al_tl[i, npy - 2] - v_tl[i, npy - 2]
                bl[i, npy - 2] = al[i, npy - 2] - v[i, npy - 2]
                xt_tl = c1 * v_tl[i, npy - 3] + c2 * v_tl[i, npy - 2] + c3 * v_tl[i, npy - 1]
                xt = c1 * v[i, npy - 3] + c2 * v[i, npy - 2] + c3 * v[i, npy - 1]
                br_tl[i, npy - 2] = xt_tl - v_tl[i, npy - 2]
                br[i, npy - 2] = xt - v[i, npy - 2]
                bl_tl[i, npy - 1] = xt_tl - v_tl[i, npy - 1]
                bl[i, npy - 1] = xt - v[i, npy - 1]
                xt_tl = 0.5 * (((2. * dy[i, npy - 1] + dy[i, npy - 2]) * v_tl[i, npy - 1] - dy[i, npy - 1] * v_tl[i, npy - 2])
                               / (dy[i, npy - 1] + dy[i, npy - 2])
                               + ((2. * dy[i, npy] + dy[i, npy + 1]) * v_tl[i, npy] - dy[i, npy] * v_tl[i, npy + 1])
                               / (dy[i, npy] + dy[i, npy + 1]))
                xt = 0.5 * (((2. * dy[i, npy - 1] + dy[i, npy - 2]) * v[i, npy - 1] - dy[i, npy - 1] * v[i, npy - 2])
                            / (dy[i, npy - 1] + dy[i, npy - 2])
                            + ((2. * dy[i, npy] + dy[i, npy + 1]) * v[i, npy] - dy[i, npy] * v[i, npy + 1])
                            / (dy[i, npy] + dy[i, npy + 1]))
                br_tl[i, npy - 1] = xt_tl - v_tl[i, npy - 1]
                br[i, npy - 1] = xt - v[i, npy - 1]
                bl_tl[i, npy] = xt_tl - v_tl[i, npy]
                bl[i, npy] = xt - v[i, npy]
                br_tl[i, npy] = c3 * v_tl[i, npy] + c2 * v_tl[i, npy + 1] + c1 * v_tl[i, npy + 2] - v_tl[i, npy]
                br[i, npy] = c3 * v[i, npy] + c2 * v[i, npy + 1] + c1 * v[i, npy + 2] - v[i, npy]
            if is_ == 1:
                for jj in (npy - 1, npy):
                    bl_tl[1, jj] = 0.0; bl[1, jj] = 0.
                    br_tl[1, jj] = 0.0; br[1, jj] = 0.
            if ie + 1 == npx:
                for jj in (npy - 1, npy):
                    bl_tl[npx, jj] = 0.0; bl[npx, jj] = 0.
                    br_tl[npx, jj] = 0.0; br[npx, jj] = 0.
        for j in range(js - 1, je + 2):
            for i in range(is_, ie + 2):
                b0_tl[i, j] = bl_tl[i, j] + br_tl[i, j]
                b0[i, j] = bl[i, j] + br[i, j]
        if jord == 2:
            for j in range(js, je + 2):
                for i in range(is_, ie + 2):
                    if c[i, j] > 0.:
                        cfl_tl = rdy[i, j - 1] * c_tl[i, j]
                        cfl = c[i, j] * rdy[i, j - 1]
                        flux_tl[i, j] = (v_tl[i, j - 1] + (1. - cfl) * (br_tl[i, j - 1] - cfl_tl * b0[i, j - 1] - cfl * b0_tl[i, j - 1])
                                         - cfl_tl * (br[i, j - 1] - cfl * b0[i, j - 1]))
                        flux[i, j] = v[i, j - 1] + (1. - cfl) * (br[i, j - 1] - cfl * b0[i, j - 1])
                    else:
                        cfl_tl = rdy[i, j] * c_tl[i, j]
                        cfl = c[i, j] * rdy[i, j]
                        flux_tl[i, j] = (v_tl[i, j] + cfl_tl * (bl[i, j] + cfl * b0[i, j])
                                         + (1. + cfl) * (bl_tl[i, j] + cfl_tl * b0[i, j] + cfl * b0_tl[i, j]))
                        flux[i, j] = v[i, j] + (1. + cfl) * (bl[i, j] + cfl * b0[i, j])
    return flux, flux_tl
