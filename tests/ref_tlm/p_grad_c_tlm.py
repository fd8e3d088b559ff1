"""P_GRAD_C_TLM (model_tlmadm/dyn_core_tlm.F90:3194-3275), transliterated for a whole cube tile."""
from . import F


def p_grad_c_tlm(dt2, npz, delpc, delpc_tl, pkc, pkc_tl, gz, gz_tl, uc, uc_tl, vc, vc_tl, bd, rdxc, rdyc, hydrostatic):
    """uc, vc (and _tl) updated in place"""
    is_, ie, js, je = bd.is_, bd.ie, bd.js, bd.je
    wk = F((is_ - 1, ie + 1), (js - 1, je + 1)); wk_tl = F((is_ - 1, ie + 1), (js - 1, je + 1))
    for k in range(1, npz + 1):
        if hydrostatic:
            for j in range(js - 1, je + 2):
                for i in range(is_ - 1, ie + 2):
                    wk_tl[i, j] = pkc_tl[i, j, k + 1] - pkc_tl[i, j, k]
                    wk[i, j] = pkc[i, j, k + 1] - pkc[i, j, k]
        else:
            for j in range(js - 1, je + 2):
                for i in range(is_ - 1, ie + 2):
                    wk_tl[i, j] = delpc_tl[i, j, k]
                    wk[i, j] = delpc[i, j, k]
        for j in range(js, je + 1):
            for i in range(is_, ie + 2):
                uc_tl[i, j, k] = (uc_tl[i, j, k] + dt2 * rdxc[i, j] * (
                    (gz_tl[i - 1, j, k + 1] - gz_tl[i, j, k]) * (pkc[i, j, k + 1] - pkc[i - 1, j, k])
                    + (gz[i - 1, j, k + 1] - gz[i, j, k]) * (pkc_tl[i, j, k + 1] - pkc_tl[i - 1, j, k])
                    + (gz_tl[i - 1, j, k] - gz_tl[i, j, k + 1]) * (pkc[i - 1, j, k + 1] - pkc[i, j, k])
                    + (gz[i - 1, j, k] - gz[i, j, k + 1]) * (pkc_tl[i - 1, j, k + 1] - pkc_tl[i, j, k])) / (wk[i - 1, j] + wk[i, j])
                    - dt2 * rdxc[i, j] * (wk_tl[i - 1, j] + wk_tl[i, j]) * (
                        (gz[i - 1, j, k + 1] - gz[i, j, k]) * (pkc[i, j, k + 1] - pkc[i - 1, j, k])
                        + (gz[i - 1, j, k] - gz[i, j, k + 1]) * (pkc[i - 1, j, k + 1] - pkc[i, j, k])) / (wk[i - 1, j] + wk[i, j]) ** 2)
                uc[i, j, k] = uc[i, j, k] + dt2 * rdxc[i, j] / (wk[i - 1, j] + wk[i, j]) * (
                    (gz[i - 1, j, k + 1] - gz[i, j, k]) * (pkc[i, j, k + 1] - pkc[i - 1, j, k])
                    + (gz[i - 1, j, k] - gz[i, j, k + 1]) * (pkc[i - 1, j, k + 1] - pkc[i, j, k]))
        for j in range(js, je + 2):
            for i in range(is_, ie + 1):
                vc_tl[i, j, k] = (vc_tl[i, j, k] + dt2 * rdyc[i, j] * (
                    (gz_tl[i, j - 1, k + 1] - gz_tl[i, j, k]) * (pkc[i, j, k + 1] - pkc[i, j - 1, k])
                    + (gz[i, j - 1, k + 1] - gz[i, j, k]) * (pkc_tl[i, j, k + 1] - pkc_tl[i, j - 1, k])
                    + (gz_tl[i, j - 1, k] - gz_tl[i, j, k + 1]) * (pkc[i, j - 1, k + 1] - pkc[i, j, k])
                    + (gz[i, j - 1, k] - gz[i, j, k + 1]) * (pkc_tl[i, j - 1, k + 1] - pkc_tl[i, j, k])) / (wk[i, j - 1] + wk[i, j])
                    - dt2 * rdyc[i, j] * (wk_tl[i, j - 1] + wk_tl[i, j]) * (
                        (gz[i, j - 1, k + 1] - gz[i, j, k]) * (pkc[i, j, k + 1] - pkc[i, j - 1, k])
                        + (gz[i, j - 1, k] - gz[i, j, k + 1]) * (pkc[i, j - 1, k + 1] - pkc[i, j, k])) / (wk[i, j - 1] + wk[i, j]) ** 2)
                vc[i, j, k] = vc[i, j, k] + dt2 * rdyc[i, j] / (wk[i, j - 1] + wk[i, j]) * (
                    (gz[i, j - 1, k + 1] - gz[i, j, k]) * (pkc[i, j, k + 1] - pkc[i, j - 1, k])
                    + (gz[i, j - 1, k] - gz[i, j, k + 1]) * (pkc[i, j - 1, k + 1] - pkc[i, j, k]))
