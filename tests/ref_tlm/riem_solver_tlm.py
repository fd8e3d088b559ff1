"""RIEM_SOLVER_C_TLM (model_tlmadm/nh_utils_tlm.F90:723-845) and RIEM_SOLVER3_TLM (model_tlmadm/nh_core_tlm.F90:49-243), transliterated in
the style of sim1_solver_tlm.py: the `DO i` loops (and the independent `DO j` loop around them) are numpy vector statements over the
columns, every `DO k` loop and the statement order are the source's.  a_imp > 0.999 (SIM1 solver), use_logp = .false.
Arrays are [km(+1), ncol], Fortran level k at row k - 1."""
import numpy as np
from .sim1_solver_tlm import sim1_solver_tlm

# grav and rdgas come from FMS constants_mod (not in the reference tree): they are arguments here


def riem_solver_c_tlm(grav, rdgas, dt, km, akap, ptop, hs, w3, w3_tl, pt, pt_tl, delp, delp_tl, gz, gz_tl, ws, ws_tl, p_fac):
    """returns gz, gz_tl, pef, pef_tl (km+1 rows)"""
    K = lambda k: k - 1
    ni = delp.shape[1]
    z = lambda n: np.zeros((n, ni))
    gz = gz.copy(); gz_tl = gz_tl.copy()
    dm, dz2, w2, pm2 = z(km), z(km), z(km), z(km)
    dm_tl, dz2_tl, w2_tl, pm2_tl = z(km), z(km), z(km), z(km)
    pem, pem_tl, pef, pef_tl = z(km + 1), z(km + 1), z(km + 1), z(km + 1)
    gama = 1. / (1. - akap)
    rgrav = 1. / grav
    for k in range(1, km + 1):
        dm_tl[K(k)] = delp_tl[K(k)]
        dm[K(k)] = delp[K(k)]
    pef_tl[K(1)] = 0.0
    pef[K(1)] = ptop
    pem_tl[K(1)] = 0.0
    pem[K(1)] = ptop
    for k in range(2, km + 2):
        pem_tl[K(k)] = pem_tl[K(k - 1)] + dm_tl[K(k - 1)]
        pem[K(k)] = pem[K(k - 1)] + dm[K(k - 1)]
    for k in range(1, km + 1):
        dz2_tl[K(k)] = gz_tl[K(k + 1)] - gz_tl[K(k)]
        dz2[K(k)] = gz[K(k + 1)] - gz[K(k)]
        arg1_tl = (pem_tl[K(k + 1)] * pem[K(k)] - pem[K(k + 1)] * pem_tl[K(k)]) / pem[K(k)] ** 2
        arg1 = pem[K(k + 1)] / pem[K(k)]
        pm2_tl[K(k)] = (dm_tl[K(k)] * np.log(arg1) - dm[K(k)] * arg1_tl / arg1) / np.log(arg1) ** 2
        pm2[K(k)] = dm[K(k)] / np.log(arg1)
        dm_tl[K(k)] = rgrav * dm_tl[K(k)]
        dm[K(k)] = dm[K(k)] * rgrav
        w2_tl[K(k)] = w3_tl[K(k)]
        w2[K(k)] = w3[K(k)]
    pe2, pe2_tl, w2, w2_tl, dz2, dz2_tl = sim1_solver_tlm(dt, km, rdgas, gama, akap, dm, dm_tl, pm2, pm2_tl, pem, pem_tl, w2, w2_tl,
                                                          dz2, dz2_tl, pt, pt_tl, ws, ws_tl, p_fac)
    for k in range(2, km + 2):
        pef_tl[K(k)] = pe2_tl[K(k)] + pem_tl[K(k)]
        pef[K(k)] = pe2[K(k)] + pem[K(k)]
    gz_tl[K(km + 1)] = 0.0
    gz[K(km + 1)] = hs
    for k in range(km, 0, -1):
        gz_tl[K(k)] = gz_tl[K(k + 1)] - grav * dz2_tl[K(k)]
        gz[K(k)] = gz[K(k + 1)] - dz2[K(k)] * grav
    return gz, gz_tl, pef, pef_tl


def riem_solver3_tlm(grav, rdgas, dt, km, akap, ptop, zs, w, w_tl, pt, pt_tl, delp, delp_tl, zh, zh_tl, ws, ws_tl, p_fac, fp_out=False):
    """returns dict of w, delz, zh, ppe, pk3, pe (= pem, the last_call output), peln and their _tl"""
    K = lambda k: k - 1
    ni = delp.shape[1]
    z = lambda n: np.zeros((n, ni))
    zh = zh.copy(); zh_tl = zh_tl.copy()
    dm, dz2, w2, pm2 = z(km), z(km), z(km), z(km)
    dm_tl, dz2_tl, w2_tl, pm2_tl = z(km), z(km), z(km), z(km)
    pem, pem_tl, peln2, peln2_tl, pk3, pk3_tl = z(km + 1), z(km + 1), z(km + 1), z(km + 1), z(km + 1), z(km + 1)
    gama = 1. / (1. - akap)
    rgrav = 1. / grav
    peln1 = np.log(ptop)
    ptk = np.exp(akap * peln1)
    for k in range(1, km + 1):
        dm_tl[K(k)] = delp_tl[K(k)]
        dm[K(k)] = delp[K(k)]
    pem_tl[K(1)] = 0.0
    pem[K(1)] = ptop
    peln2_tl[K(1)] = 0.0
    peln2[K(1)] = peln1
    pk3_tl[K(1)] = 0.0
    pk3[K(1)] = ptk
    for k in range(2, km + 2):
        pem_tl[K(k)] = pem_tl[K(k - 1)] + dm_tl[K(k - 1)]
        pem[K(k)] = pem[K(k - 1)] + dm[K(k - 1)]
        peln2_tl[K(k)] = pem_tl[K(k)] / pem[K(k)]
        peln2[K(k)] = np.log(pem[K(k)])
        pk3_tl[K(k)] = akap * peln2_tl[K(k)] * np.exp(akap * peln2[K(k)])
        pk3[K(k)] = np.exp(akap * peln2[K(k)])
    for k in range(1, km + 1):
        pm2_tl[K(k)] = ((dm_tl[K(k)] * (peln2[K(k + 1)] - peln2[K(k)]) - dm[K(k)] * (peln2_tl[K(k + 1)] - peln2_tl[K(k)]))
                        / (peln2[K(k + 1)] - peln2[K(k)]) ** 2)
        pm2[K(k)] = dm[K(k)] / (peln2[K(k + 1)] - peln2[K(k)])
        dm_tl[K(k)] = rgrav * dm_tl[K(k)]
        dm[K(k)] = dm[K(k)] * rgrav
        dz2_tl[K(k)] = zh_tl[K(k + 1)] - zh_tl[K(k)]
        dz2[K(k)] = zh[K(k + 1)] - zh[K(k)]
        w2_tl[K(k)] = w_tl[K(k)]
        w2[K(k)] = w[K(k)]
    pe2, pe2_tl, w2, w2_tl, dz2, dz2_tl = sim1_solver_tlm(dt, km, rdgas, gama, akap, dm, dm_tl, pm2, pm2_tl, pem, pem_tl, w2, w2_tl,
                                                          dz2, dz2_tl, pt, pt_tl, ws, ws_tl, p_fac)
    out = dict(w=w2, w_tl=w2_tl, delz=dz2, delz_tl=dz2_tl, peln=peln2, peln_tl=peln2_tl, pk3=pk3, pk3_tl=pk3_tl, pe=pem, pe_tl=pem_tl)
    if fp_out:
        out["ppe_tl"] = pe2_tl + pem_tl
        out["ppe"] = pe2 + pem
    else:
        out["ppe_tl"] = pe2_tl
        out["ppe"] = pe2
    zh_tl[K(km + 1)] = 0.0
    zh[K(km + 1)] = zs
    for k in range(km, 0, -1):
        zh_tl[K(k)] = zh_tl[K(k + 1)] - dz2_tl[K(k)]
        zh[K(k)] = zh[K(k + 1)] - dz2[K(k)]
    out["zh"] = zh; out["zh_tl"] = zh_tl
    return out
