"""ctypes binding of the C ABI in include/fv3lm_b200.h (used by tests/ and bench.py; a
Fortran host binds the same symbols through ISO_C_BINDING).

There is no CPU fallback: `load()` opens libfv3lm_b200.so (nvcc, sm_100a) and every
compute entry point fails loudly without a GPU.  `load(emu=True)` opens the TEST-ONLY
host emulation build of the same stage functors (tests marked "not gpu" use it to check
stage arithmetic against the oracle); nothing in the product path ever asks for it.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

MODE_NL, MODE_TL, MODE_AD = 0, 1, 2


class TrajFlags(C.Structure):
    """the nonlinear model's switches in two-sided mode (fv3lm_config.traj)"""
    _fields_ = [("hord_mt", C.c_int), ("hord_vt", C.c_int), ("hord_tm", C.c_int), ("hord_dp", C.c_int), ("hord_tr", C.c_int),
                ("nord", C.c_int), ("do_vort_damp", C.c_int), ("n_sponge", C.c_int),
                ("kord_mt", C.c_int), ("kord_wz", C.c_int), ("kord_tm", C.c_int), ("kord_tr", C.c_int),
                ("dddmp", C.c_double), ("d2_bg", C.c_double), ("d4_bg", C.c_double), ("vtdm4", C.c_double),
                ("d2_bg_k1", C.c_double), ("d2_bg_k2", C.c_double)]


class Config(C.Structure):
    _fields_ = [("npx", C.c_int), ("npy", C.c_int), ("npz", C.c_int), ("ng", C.c_int), ("ntiles", C.c_int),
                ("hydrostatic", C.c_int), ("n_split", C.c_int), ("k_split", C.c_int), ("nq", C.c_int),
                ("hord_mt", C.c_int), ("hord_vt", C.c_int), ("hord_tm", C.c_int), ("hord_dp", C.c_int),
                ("hord_tr", C.c_int), ("n_sponge", C.c_int), ("nord", C.c_int),
                ("dt", C.c_double), ("ptop", C.c_double),
                ("dddmp", C.c_double), ("d2_bg", C.c_double), ("d4_bg", C.c_double), ("vtdm4", C.c_double),
                ("d2_bg_k1", C.c_double), ("d2_bg_k2", C.c_double), ("d_ext", C.c_double), ("beta", C.c_double),
                ("zvir", C.c_double), ("kappa", C.c_double), ("cp", C.c_double), ("rdgas", C.c_double),
                ("grav", C.c_double), ("do_vort_damp", C.c_int),
                ("rank", C.c_int), ("nranks", C.c_int), ("layout_x", C.c_int), ("layout_y", C.c_int), ("device", C.c_int),
                ("a_imp", C.c_double), ("p_fac", C.c_double), ("d_con", C.c_double),
                ("two_sided", C.c_int), ("split_damp", C.c_int), ("hord_ks_pert", C.c_int), ("hord_ks_traj", C.c_int),
                ("q_split_dynamic", C.c_int), ("q_split_max", C.c_int), ("traj", TrajFlags), ("d2_bg_ks", C.c_double)]


class Fields(C.Structure):
    _fields_ = [(n, C.POINTER(C.c_double)) for n in ("u", "v", "t", "delp", "qv", "ql", "qi", "o3", "w", "delz")]


EXPORTS = ["fv3lm_decomp_info", "fv3lm_nccl_unique_id", "fv3lm_comm_init_nccl", "fv3lm_comm_set_callback", "fv3lm_comm_stats",
           "fv3lm_profile_steps", "fv3lm_set_phis", "fv3lm_traj_set", "fv3lm_traj_get", "fv3lm_step_nl", "fv3lm_step_tl", "fv3lm_step_ad",
           "fv3lm_pert_upload", "fv3lm_pert_download", "fv3lm_step_tl_dev", "fv3lm_step_ad_dev", "fv3lm_time_steps",
           "fv3lm_program_stats", "fv3lm_create", "fv3lm_destroy", "fv3lm_last_error", "fv3lm_set_metric", "fv3lm_set_metric_scalar",
           "fv3lm_module_run", "fv3lm_module_list", "fv3lm_launch_count", "fv3lm_pool_peak_bytes", "fv3lm_sync",
           "fv3lm_turb_set_ltraj", "fv3lm_turb_step_nl", "fv3lm_turb_step_tl", "fv3lm_turb_step_ad", "fv3lm_turb_step_tl_dev",
           "fv3lm_turb_step_ad_dev", "fv3lm_time_turb", "fv3lm_set_c2l", "fv3lm_traj_get_winds"]

TURB_ARRAYS = ["akv", "bkv", "ckv", "aks", "bks", "cks", "akq", "bkq", "ckq", "pk"]


class TurbCoeffs(C.Structure):
    """mirror of fv3lm_turb_coeffs (include/fv3lm_b200.h)"""
    _fields_ = [(n, C.POINTER(C.c_double)) for n in TURB_ARRAYS] + [("decomposed", C.c_int)]

METRICS_2D = ["area", "rarea", "area_c", "rarea_c", "dx", "dy", "rdx", "rdy", "dxa", "dya", "rdxa", "rdya", "dxc",
              "dyc", "rdxc", "rdyc", "cosa", "sina", "rsina", "cosa_u", "sina_u", "rsin_u", "cosa_v", "sina_v",
              "rsin_v", "cosa_s", "rsin2", "divg_u", "divg_v", "del6_u", "del6_v", "f0", "fC"]
METRICS_1D = ["edge_w", "edge_e", "edge_s", "edge_n", "edge_vect_w", "edge_vect_e", "edge_vect_s", "edge_vect_n"]


def lib_path(emu=False):
    return os.path.join(_HERE, "libfv3lm_hostemu.so" if emu else "libfv3lm_b200.so")


def load(emu=False):
    path = lib_path(emu)
    if not os.path.exists(path):
        raise RuntimeError("%s is missing: run __graft_entry__.build() (there is no CPU fallback)" % path)
    lib = C.CDLL(path)
    lib.fv3lm_last_error.restype = C.c_char_p
    lib.fv3lm_module_list.restype = C.c_char_p
    lib.fv3lm_launch_count.restype = C.c_longlong
    lib.fv3lm_pool_peak_bytes.restype = C.c_double
    return lib


def default_config(N, npz, **kw):
    """defaults follow SURVEY 8(d): linear schemes, split_* = false"""
    rdgas = 8314.47 / 28.965
    cfg = Config()
    cfg.npx = cfg.npy = N + 1
    cfg.npz = npz; cfg.ng = 3; cfg.ntiles = 6
    cfg.hydrostatic = 1; cfg.n_split = 1; cfg.k_split = 1; cfg.nq = 4
    cfg.hord_mt = cfg.hord_vt = cfg.hord_tm = cfg.hord_dp = cfg.hord_tr = 2
    cfg.n_sponge = 0; cfg.nord = 1
    cfg.dt = 900.0; cfg.ptop = 1.0
    cfg.dddmp = 0.2; cfg.d2_bg = 0.015; cfg.d4_bg = 0.15; cfg.vtdm4 = 0.0005
    cfg.d2_bg_k1 = 4.0; cfg.d2_bg_k2 = 2.0; cfg.d_ext = 0.0; cfg.beta = 0.0     # d_ext > 0 selects a path that is not built: fv3lm_create refuses it
    cfg.rdgas = rdgas; cfg.cp = 3.5 * rdgas; cfg.kappa = rdgas / (3.5 * rdgas)
    cfg.zvir = (8314.47 / 18.015) / rdgas - 1.0; cfg.grav = 9.80665
    cfg.do_vort_damp = 1
    for k, v in kw.items():
        if k == "traj":                       # dict of the nonlinear model's switches -> two-sided mode
            cfg.two_sided = 1
            for kk, vv in v.items():
                setattr(cfg.traj, kk, vv)
        else:
            setattr(cfg, k, v)
    return cfg


class FV3LM:
    """One handle = one GPU's share of the cubed sphere (whole sphere for a single GPU)."""

    def __init__(self, cfg, ak=None, bk=None, emu=False):
        self.lib = load(emu)
        self.cfg = cfg
        self.h = C.c_void_p()
        akp = bkp = None
        if ak is not None:
            self._ak = np.ascontiguousarray(ak, dtype=np.float64); self._bk = np.ascontiguousarray(bk, dtype=np.float64)
            akp = self._ak.ctypes.data_as(C.POINTER(C.c_double)); bkp = self._bk.ctypes.data_as(C.POINTER(C.c_double))
        rc = self.lib.fv3lm_create(C.byref(cfg), akp, bkp, C.byref(self.h))
        if rc != 0:
            raise RuntimeError("fv3lm_create failed: %s" % self.lib.fv3lm_last_error(None).decode())
        self.N = cfg.npx - 1
        self.NX = self.N + 2 * cfg.ng + 1
        info = (C.c_int * 6)(); t = (C.c_int * 64)(); i0 = (C.c_int * 64)(); j0 = (C.c_int * 64)()
        self.lib.fv3lm_decomp_info(self.h, info, t, i0, j0)
        self.nsub, self.nxl, self.nyl, self.lx, self.ly, self.nsub_total = [int(x) for x in info]
        self.sub_tile = [int(t[l]) for l in range(self.nsub)]
        self.sub_i0 = [int(i0[l]) for l in range(self.nsub)]
        self.sub_j0 = [int(j0[l]) for l in range(self.nsub)]
        self.NXl = self.nxl + 2 * cfg.ng + 1
        self.NYl = self.nyl + 2 * cfg.ng + 1
        self.whole = (self.nsub == 6 and self.nxl == self.N and self.nyl == self.N)
        self._cb = None

    # ---- global (whole cube, [6, ..., N+7, N+7]) <-> this rank's sub-domain arrays ----------------
    def scatter(self, a):
        """halo'd global array [6, nk, NYg, NXg] -> [nsub, nk, NYl, NXl] (every cell the sub-domain array holds)"""
        if self.whole:
            return np.ascontiguousarray(a)
        return np.ascontiguousarray(np.stack([a[t, ..., j0:j0 + self.NYl, i0:i0 + self.NXl]
                                              for t, i0, j0 in zip(self.sub_tile, self.sub_i0, self.sub_j0)]))

    def scatter_owned(self, a):
        """like scatter, but only the cells a sub-domain owns (1..nxl, 1..nyl); zero elsewhere"""
        o = self.cfg.ng - 1
        out = np.zeros((self.nsub,) + a.shape[1:-2] + (self.NYl, self.NXl))
        for l, (t, i0, j0) in enumerate(zip(self.sub_tile, self.sub_i0, self.sub_j0)):
            out[l, ..., o + 1:o + 1 + self.nyl, o + 1:o + 1 + self.nxl] = a[t, ..., j0 + o + 1:j0 + o + 1 + self.nyl, i0 + o + 1:i0 + o + 1 + self.nxl]
        return out

    def gather(self, loc, out, closed=True):
        """write the (closed: + shared edge row/column) compute region of each sub-domain into the global array"""
        o = self.cfg.ng - 1; e = 1 if closed else 0
        for l, (t, i0, j0) in enumerate(zip(self.sub_tile, self.sub_i0, self.sub_j0)):
            out[t, ..., j0 + o + 1:j0 + o + 1 + self.nyl + e, i0 + o + 1:i0 + o + 1 + self.nxl + e] = \
                loc[l, ..., o + 1:o + 1 + self.nyl + e, o + 1:o + 1 + self.nxl + e]
        return out

    def gather_add(self, loc, out):
        """sum every copy of every cell (adjoint of scatter)"""
        for l, (t, i0, j0) in enumerate(zip(self.sub_tile, self.sub_i0, self.sub_j0)):
            out[t, ..., j0:j0 + self.NYl, i0:i0 + self.NXl] += loc[l]
        return out

    def scatter_c(self, a):
        """compute-domain global array [6, nk, N, N] -> [nsub, nk, nyl, nxl]"""
        if self.whole:
            return np.ascontiguousarray(a)
        return np.ascontiguousarray(np.stack([a[t, ..., j0:j0 + self.nyl, i0:i0 + self.nxl]
                                              for t, i0, j0 in zip(self.sub_tile, self.sub_i0, self.sub_j0)]))

    def gather_c(self, loc, out):
        for l, (t, i0, j0) in enumerate(zip(self.sub_tile, self.sub_i0, self.sub_j0)):
            out[t, ..., j0:j0 + self.nyl, i0:i0 + self.nxl] = loc[l]
        return out

    # ---- inter-rank transport ---------------------------------------------------------------------
    def comm_init_nccl(self, id128):
        self._check(self.lib.fv3lm_comm_init_nccl(self.h, C.c_char_p(bytes(id128))), "comm_init_nccl")

    def nccl_unique_id(self):
        buf = C.create_string_buffer(128)
        if self.lib.fv3lm_nccl_unique_id(buf) != 0:
            raise RuntimeError("fv3lm_nccl_unique_id failed: %s" % self.lib.fv3lm_last_error(None).decode())
        return buf.raw

    def comm_set_callback(self, fn):
        """TEST-ONLY (host emulation): fn(peers, sendbufs, recvbufs) with numpy views"""
        CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_size_t),
                         C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_size_t))

        def tramp(user, n, peers, sb, sc, rb, rc):
            ps = [int(peers[k]) for k in range(n)]
            sends = [np.ctypeslib.as_array(sb[k], shape=(int(sc[k]),)) if sc[k] else np.zeros(0) for k in range(n)]
            recvs = [np.ctypeslib.as_array(rb[k], shape=(int(rc[k]),)) if rc[k] else np.zeros(0) for k in range(n)]
            fn(ps, sends, recvs)
        self._cb = CB(tramp)
        self._check(self.lib.fv3lm_comm_set_callback(self.h, self._cb, None), "comm_set_callback")

    def comm_stats(self):
        out = (C.c_double * 2)()
        self.lib.fv3lm_comm_stats(self.h, out)
        return int(out[0]), float(out[1])

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError("%s failed: %s" % (what, self.lib.fv3lm_last_error(self.h).decode()))

    def set_metrics(self, M):
        """M: dict from oracle.grid.build_metrics (or the host model's gridstruct)"""
        dp = C.POINTER(C.c_double)
        def up(name, arr, is1d):
            if is1d:
                a = np.ascontiguousarray(np.asarray(arr, dtype=np.float64)[self.sub_tile])
            else:
                a = np.ascontiguousarray(self.scatter(np.asarray(arr, dtype=np.float64)), dtype=np.float64)
            self._check(self.lib.fv3lm_set_metric(self.h, name.encode(), a.ctypes.data_as(dp), int(is1d)), "set_metric " + name)
        for n in METRICS_2D:
            up(n, M[n], False)
        for k in (1, 2, 3, 4):
            up("sin_sg%d" % k, M["sin_sg"][..., k], False)
            up("cos_sg%d" % k, M["cos_sg"][..., k], False)
        up("agrid_lon", M["agrid"][..., 0], False); up("agrid_lat", M["agrid"][..., 1], False)
        up("grid_lon", M["grid"][..., 0], False); up("grid_lat", M["grid"][..., 1], False)
        for n in METRICS_1D:
            up(n, M[n], True)
        for n in ("da_min", "da_min_c"):
            self._check(self.lib.fv3lm_set_metric_scalar(self.h, n.encode(), C.c_double(M[n])), "set_metric_scalar")

    def module_run(self, module, mode, traj, pert=None, params=None):
        """traj / pert: dict name -> float64 ndarray [6, nk, NY, NX] (modified in place)."""
        pert = pert or {}
        params = params or {}
        names = list(traj.keys())
        for n in pert:
            if n not in traj:
                raise KeyError("pert field %s has no traj array" % n)
        dp = C.POINTER(C.c_double)
        n = len(names)
        for k in names:
            a = traj[k]
            assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], k
        c_names = (C.c_char_p * n)(*[s.encode() for s in names])
        c_traj = (dp * n)(*[traj[k].ctypes.data_as(dp) for k in names])
        c_pert = (dp * n)(*[(pert[k].ctypes.data_as(dp) if k in pert else C.cast(None, dp)) for k in names])
        pn = list(params.keys())
        c_pn = (C.c_char_p * max(1, len(pn)))(*[s.encode() for s in pn]) if pn else (C.c_char_p * 1)()
        c_pv = (C.c_double * max(1, len(pn)))(*[float(params[k]) for k in pn]) if pn else (C.c_double * 1)()
        rc = self.lib.fv3lm_module_run(self.h, module.encode(), int(mode), n, c_names, c_traj, c_pert, len(pn), c_pn, c_pv)
        self._check(rc, "module_run(%s)" % module)

    # ---- step-level API (mirror of fv3jedi_lm_dynamics_mod step_nl / step_tl / step_ad) ----
    FIELDS = ["u", "v", "t", "delp", "qv", "ql", "qi", "o3", "w", "delz"]

    def _fields(self, d):
        """d: dict name -> float64 ndarray [6, K, N, N] (compute domain).  Returns (struct, keepalive)"""
        dp = C.POINTER(C.c_double)
        st = Fields()
        keep = []
        for n in self.FIELDS:
            if n in d:
                a = d[n]
                assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], n
                setattr(st, n, a.ctypes.data_as(dp)); keep.append(a)
            else:
                setattr(st, n, C.cast(None, dp))
        return st, keep

    def set_phis(self, phis):
        a = np.ascontiguousarray(phis, dtype=np.float64)
        self._check(self.lib.fv3lm_set_phis(self.h, a.ctypes.data_as(C.POINTER(C.c_double))), "set_phis")

    def traj_set(self, slot, traj):
        st, _k = self._fields(traj)
        self._check(self.lib.fv3lm_traj_set(self.h, int(slot), C.byref(st)), "traj_set")

    def traj_get(self, slot, traj):
        st, _k = self._fields(traj)
        self._check(self.lib.fv3lm_traj_get(self.h, int(slot), C.byref(st)), "traj_get")

    def step_nl(self, slot_in, slot_out):
        self._check(self.lib.fv3lm_step_nl(self.h, int(slot_in), int(slot_out)), "step_nl")

    def set_c2l(self, a11, a12, a21, a22):
        """gridstruct%a11 .. a22 on the compute domain of this handle's sub-domains ([nsub, nyl, nxl])"""
        dp = C.POINTER(C.c_double)
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (a11, a12, a21, a22)]
        self._check(self.lib.fv3lm_set_c2l(self.h, *[a.ctypes.data_as(dp) for a in arrs]), "set_c2l")

    def traj_get_winds(self, ua, va):
        dp = C.POINTER(C.c_double)
        assert ua.dtype == np.float64 and va.dtype == np.float64 and ua.flags["C_CONTIGUOUS"] and va.flags["C_CONTIGUOUS"]
        self._check(self.lib.fv3lm_traj_get_winds(self.h, ua.ctypes.data_as(dp), va.ctypes.data_as(dp)), "traj_get_winds")

    def step_tl(self, slot, pert):
        st, _k = self._fields(pert)
        self._check(self.lib.fv3lm_step_tl(self.h, int(slot), C.byref(st)), "step_tl")

    def step_ad(self, slot, pert):
        st, _k = self._fields(pert)
        self._check(self.lib.fv3lm_step_ad(self.h, int(slot), C.byref(st)), "step_ad")

    def pert_upload(self, pert):
        st, _k = self._fields(pert)
        self._check(self.lib.fv3lm_pert_upload(self.h, C.byref(st)), "pert_upload")

    def pert_download(self, pert):
        st, _k = self._fields(pert)
        self._check(self.lib.fv3lm_pert_download(self.h, C.byref(st)), "pert_download")

    def step_tl_dev(self, slot):
        self._check(self.lib.fv3lm_step_tl_dev(self.h, int(slot)), "step_tl_dev")

    def step_ad_dev(self, slot):
        self._check(self.lib.fv3lm_step_ad_dev(self.h, int(slot)), "step_ad_dev")

    # ---- linearised boundary-layer turbulence (fv3jedi_lm_turbulence_mod.F90) ----
    def turb_set_ltraj(self, slot, coeffs, decomposed=False):
        """coeffs: dict akv bkv ckv aks bks cks akq bkq ckq [pk] of compact float64 arrays (BL_DRIVER's diagonals)"""
        st = TurbCoeffs()
        keep = []
        for n in TURB_ARRAYS:
            a = coeffs.get(n)
            if a is None:
                continue
            a = np.ascontiguousarray(a, dtype=np.float64)
            keep.append(a)
            setattr(st, n, a.ctypes.data_as(C.POINTER(C.c_double)))
        st.decomposed = int(decomposed)
        self._check(self.lib.fv3lm_turb_set_ltraj(self.h, int(slot), C.byref(st)), "turb_set_ltraj")

    def turb_step_nl(self, slot_ltraj, slot_state=None):
        slot_state = slot_ltraj if slot_state is None else slot_state
        self._check(self.lib.fv3lm_turb_step_nl(self.h, int(slot_ltraj), int(slot_state)), "turb_step_nl")

    def turb_step_tl(self, slot, pert):
        st, _k = self._fields(pert)
        self._check(self.lib.fv3lm_turb_step_tl(self.h, int(slot), C.byref(st)), "turb_step_tl")

    def turb_step_ad(self, slot, pert):
        st, _k = self._fields(pert)
        self._check(self.lib.fv3lm_turb_step_ad(self.h, int(slot), C.byref(st)), "turb_step_ad")

    def turb_step_tl_dev(self, slot):
        self._check(self.lib.fv3lm_turb_step_tl_dev(self.h, int(slot)), "turb_step_tl_dev")

    def turb_step_ad_dev(self, slot):
        self._check(self.lib.fv3lm_turb_step_ad_dev(self.h, int(slot)), "turb_step_ad_dev")

    def time_turb(self, slot, warmup, iters):
        ms = (C.c_double * 2)()
        self._check(self.lib.fv3lm_time_turb(self.h, int(slot), int(warmup), int(iters), ms), "time_turb")
        return float(ms[0]), float(ms[1])

    def time_steps(self, slot, warmup, iters):
        ms = (C.c_double * 2)()
        self._check(self.lib.fv3lm_time_steps(self.h, int(slot), int(warmup), int(iters), ms), "time_steps")
        return float(ms[0]), float(ms[1])

    def program_stats(self, module):
        out = (C.c_double * 4)()
        self._check(self.lib.fv3lm_program_stats(self.h, module.encode(), out), "program_stats")
        return dict(ops=int(out[0]), values=int(out[1]), bytes_all_values=float(out[2]), patch_ops=int(out[3]))

    def sync(self):
        self._check(self.lib.fv3lm_sync(self.h), "sync")

    def launch_count(self):
        return int(self.lib.fv3lm_launch_count())

    def close(self):
        if self.h:
            self.lib.fv3lm_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
