"""NH_P_GRAD_TLM (model_tlmadm/dyn_core_tlm.F90:3340-3470), transliterated for a whole cube tile, use_logp = .false.; A2B_ORD4_TLM is
a2b_ord4_tlm.py (its `replace` block, a2b_edge_tlm.F90:1148-1157, is applied here)."""
from . import F
from .a2b_ord4_tlm import a2b_ord4_tlm


def _a2b_level(a, a_tl, k, a2b, bd, replace):
    """A2B_ORD4_TLM on the section a(isd:ied, jsd:jed, k); returns wk1, wk1_tl"""
    isd, ied, jsd, jed = bd.isd, bd.ied, bd.jsd, bd.jed
    qin = F((isd, ied), (jsd, jed)); qin_tl = F((isd, ied), (jsd, jed))
    for j in range(jsd, jed + 1):
        for i in range(isd, ied + 1):
            qin[i, j] = a[i, j, k]; qin_tl[i, j] = a_tl[i, j, k]
    qout, qout_tl = a2b(qin, qin_tl)
    if replace:
        for j in range(bd.js, bd.je + 2):
            for i in range(bd.is_, bd.ie + 2):
                a_tl[i, j, k] = qout_tl[i, j]
                a[i, j, k] = qout[i, j]
    return qout, qout_tl


def nh_p_grad_tlm(u, u_tl, v, v_tl, pp, pp_tl, gz, gz_tl, delp, delp_tl, pk, pk_tl, dt, rdx, rdy, bd, npz, ptk, a2b):
    """all fields F((isd,ied+1),(jsd,jed+1),(1,npz[+1])), modified in place as in the source; a2b(qin, qin_tl) -> qout, qout_tl applies
    A2B_ORD4_TLM with the tile's metrics.  ptk: the module variable (top value of pk)."""
    is_, ie, js, je = bd.is_, bd.ie, bd.js, bd.je
    wk = F((is_, ie + 1), (js, je + 1)); wk_tl = F((is_, ie + 1), (js, je + 1))
    top_value = ptk
    for j in range(js, je + 2):
        for i in range(is_, ie + 2):
            pp_tl[i, j, 1] = 0.0
            pp[i, j, 1] = 0.
            pk_tl[i, j, 1] = 0.0
            pk[i, j, 1] = top_value
    for k in range(1, npz + 2):
        if k != 1:
            _a2b_level(pp, pp_tl, k, a2b, bd, True)
            _a2b_level(pk, pk_tl, k, a2b, bd, True)
        _a2b_level(gz, gz_tl, k, a2b, bd, True)
    for k in range(1, npz + 1):
        wk1, wk1_tl = _a2b_level(delp, delp_tl, k, a2b, bd, False)
        for j in range(js, je + 2):
            for i in range(is_, ie + 2):
                wk_tl[i, j] = pk_tl[i, j, k + 1] - pk_tl[i, j, k]
                wk[i, j] = pk[i, j, k + 1] - pk[i, j, k]
        for j in range(js, je + 2):
            for i in range(is_, ie + 1):
                du1_tl = (dt * ((gz_tl[i, j, k + 1] - gz_tl[i + 1, j, k]) * (pk[i + 1, j, k + 1] - pk[i, j, k])
                                + (gz[i, j, k + 1] - gz[i + 1, j, k]) * (pk_tl[i + 1, j, k + 1] - pk_tl[i, j, k])
                                + (gz_tl[i, j, k] - gz_tl[i + 1, j, k + 1]) * (pk[i, j, k + 1] - pk[i + 1, j, k])
                                + (gz[i, j, k] - gz[i + 1, j, k + 1]) * (pk_tl[i, j, k + 1] - pk_tl[i + 1, j, k])) / (wk[i, j] + wk[i + 1, j])
                          - dt * (wk_tl[i, j] + wk_tl[i + 1, j]) * ((gz[i, j, k + 1] - gz[i + 1, j, k]) * (pk[i + 1, j, k + 1] - pk[i, j, k])
                                                                    + (gz[i, j, k] - gz[i + 1, j, k + 1]) * (pk[i, j, k + 1] - pk[i + 1, j, k]))
                          / (wk[i, j] + wk[i + 1, j]) ** 2)
                du1 = dt / (wk[i, j] + wk[i + 1, j]) * ((gz[i, j, k + 1] - gz[i + 1, j, k]) * (pk[i + 1, j, k + 1] - pk[i, j, k])
                                                        + (gz[i, j, k] - gz[i + 1, j, k + 1]) * (pk[i, j, k + 1] - pk[i + 1, j, k]))
                u_tl[i, j, k] = rdx[i, j] * (u_tl[i, j, k] + du1_tl + dt * (
                    (gz_tl[i, j, k + 1] - gz_tl[i + 1, j, k]) * (pp[i + 1, j, k + 1] - pp[i, j, k])
                    + (gz[i, j, k + 1] - gz[i + 1, j, k]) * (pp_tl[i + 1, j, k + 1] - pp_tl[i, j, k])
                    + (gz_tl[i, j, k] - gz_tl[i + 1, j, k + 1]) * (pp[i, j, k + 1] - pp[i + 1, j, k])
                    + (gz[i, j, k] - gz[i + 1, j, k + 1]) * (pp_tl[i, j, k + 1] - pp_tl[i + 1, j, k])) / (wk1[i, j] + wk1[i + 1, j])
                    - dt * (wk1_tl[i, j] + wk1_tl[i + 1, j]) * ((gz[i, j, k + 1] - gz[i + 1, j, k]) * (pp[i + 1, j, k + 1] - pp[i, j, k])
                                                                + (gz[i, j, k] - gz[i + 1, j, k + 1]) * (pp[i, j, k + 1] - pp[i + 1, j, k]))
                    / (wk1[i, j] + wk1[i + 1, j]) ** 2)
                u[i, j, k] = (u[i, j, k] + du1 + dt / (wk1[i, j] + wk1[i + 1, j]) * (
                    (gz[i, j, k + 1] - gz[i + 1, j, k]) * (pp[i + 1, j, k + 1] - pp[i, j, k])
                    + (gz[i, j, k] - gz[i + 1, j, k + 1]) * (pp[i, j, k + 1] - pp[i + 1, j, k]))) * rdx[i, j]
        for j in range(js, je + 1):
            for i in range(is_, ie + 2):
                dv1_tl = (dt * ((gz_tl[i, j, k + 1] - gz_tl[i, j + 1, k]) * (pk[i, j + 1, k + 1] - pk[i, j, k])
                                + (gz[i, j, k + 1] - gz[i, j + 1, k]) * (pk_tl[i, j + 1, k + 1] - pk_tl[i, j, k])
                                + (gz_tl[i, j, k] - gz_tl[i, j + 1, k + 1]) * (pk[i, j, k + 1] - pk[i, j + 1, k])
                                + (gz[i, j, k] - gz[i, j + 1, k + 1]) * (pk_tl[i, j, k + 1] - pk_tl[i, j + 1, k])) / (wk[i, j] + wk[i, j + 1])
                          - dt * (wk_tl[i, j] + wk_tl[i, j + 1]) * ((gz[i, j, k + 1] - gz[i, j + 1, k]) * (pk[i, j + 1, k + 1] - pk[i, j, k])
                                                                    + (gz[i, j, k] - gz[i, j + 1, k + 1]) * (pk[i, j, k + 1] - pk[i, j + 1, k]))
                          / (wk[i, j] + wk[i, j + 1]) ** 2)
                dv1 = dt / (wk[i, j] + wk[i, j + 1]) * ((gz[i, j, k + 1] - gz[i, j + 1, k]) * (pk[i, j + 1, k + 1] - pk[i, j, k])
                                                        + (gz[i, j, k] - gz[i, j + 1, k + 1]) * (pk[i, j, k + 1] - pk[i, j + 1, k]))
                v_tl[i, j, k] = rdy[i, j] * (v_tl[i, j, k] + dv1_tl + dt * (
                    (gz_tl[i, j, k + 1] - gz_tl[i, j + 1, k]) * (pp[i, j + 1, k + 1] - pp[i, j, k])
                    + (gz[i, j, k + 1] - gz[i, j + 1, k]) * (pp_tl[i, j + 1, k + 1] - pp_tl[i, j, k])
                    + (gz_tl[i, j, k] - gz_tl[i, j + 1, k + 1]) * (pp[i, j, k + 1] - pp[i, j + 1, k])
                    + (gz[i, j, k] - gz[i, j + 1, k + 1]) * (pp_tl[i, j, k + 1] - pp_tl[i, j + 1, k])) / (wk1[i, j] + wk1[i, j + 1])
                    - dt * (wk1_tl[i, j] + wk1_tl[i, j + 1]) * ((gz[i, j, k + 1] - gz[i, j + 1, k]) * (pp[i, j + 1, k + 1] - pp[i, j, k])
                                                                + (gz[i, j, k] - gz[i, j + 1, k + 1]) * (pp[i, j, k + 1] - pp[i, j + 1, k]))
                    / (wk1[i, j] + wk1[i, j + 1]) ** 2)
                v[i, j, k] = (v[i, j, k] + dv1 + dt / (wk1[i, j] + wk1[i, j + 1]) * (
                    (gz[i, j, k + 1] - gz[i, j + 1, k]) * (pp[i, j + 1, k + 1] - pp[i, j, k])
                    + (gz[i, j, k] - gz[i, j + 1, k + 1]) * (pp[i, j, k + 1] - pp[i, j + 1, k]))) * rdy[i, j]
