"""Host-side mirror of the reference's top level, src/fv3jedi_lm_mod.F90 (`fv3jedi_lm_type`): the same members (conf, traj,
pert), the same procedures (create, init_nl/tl/ad, step_nl/tl/ad, final_nl/tl/ad, delete), the same sequencing (dynamics then
physics in step_nl / step_tl :153-170, physics then dynamics in step_ad :179-185, the internal part of pert zeroed around every
TL / AD step :243-252) and the same switches (conf.do_dyn, do_phy, do_phy_trb, do_phy_mst, n, nt, saveltraj;
src/utils/fv3jedi_lm_utils_mod.F90:13-32).  Everything numerical happens behind the C ABI (fv3lm.FV3LM): increments stay on the
device between the dynamics and the physics of a step.

What the reference reads from namelists (input.nml fv_core_nml, inputpert.nml) comes in as `flags` (fv3lm_config members) and
what it takes from FMS / the grid generator as `metrics` (gridstruct arrays).  BL_DRIVER, the nonlinear boundary-layer scheme
that set_ltraj runs once per trajectory time level (src/physics/turbulence/fv3jedi_lm_turbulence_mod.F90:455-507), is the
caller's: `bl_driver(traj) -> {akv, bkv, ckv, aks, bks, cks, akq, bkq, ckq}`.  Moist physics is not built: do_phy_mst = 1 is an
error at create (the reference would run it)."""
import numpy as np
import fv3lm

PROGNOSTIC = ["u", "v", "t", "delp", "qv", "ql", "qi", "o3", "w", "delz"]


class fv3jedi_lm_conf:
    """src/utils/fv3jedi_lm_utils_mod.F90:13-32 (the members this path uses)"""

    def __init__(self):
        self.dt = 0.0
        self.saveltraj = False
        self.n = 1
        self.nt = 1
        self.ptop = 0.0
        self.npx = self.npy = self.npz = 0
        self.im = self.jm = self.lm = 0
        self.do_dyn = 1
        self.do_phy = 1
        self.do_phy_trb = 1
        self.do_phy_mst = 1
        self.ak = self.bk = None
        self.hydrostatic = True
        self.rpe = True


class fv3jedi_lm_type:
    def __init__(self, flags=None, metrics=None, bl_driver=None, emu=False):
        """flags: fv3lm_config members (hydrostatic, n_split, hord_*, traj=dict(...), rank, nranks, layout_* ...);
        metrics: dict of gridstruct arrays (fv3lm.FV3LM.set_metrics; with a11 .. a22 step_nl also returns traj['ua'], traj['va']);
        emu: TEST-ONLY host-emulation library"""
        self.conf = fv3jedi_lm_conf()
        self.traj = {}
        self.pert = {}
        self._flags = dict(flags or {})
        self._metrics = metrics
        self.bl_driver = bl_driver
        self._emu = emu
        self._h = None
        self._ltraj_set = set()
        self._winds = False

    # ---- create / delete (:43-102, :218-233) ------------------------------------------------------------------------------
    def create(self, dt, npx, npy, npz, ptop, ak, bk):
        c = self.conf
        c.dt, c.ptop = float(dt), float(ptop)
        c.ak = np.ascontiguousarray(ak, dtype=np.float64); c.bk = np.ascontiguousarray(bk, dtype=np.float64)
        if len(c.ak) != npz + 1 or len(c.bk) != npz + 1:
            raise ValueError("fv3jedi_lm create: ak, bk must hold npz + 1 values")
        if npx != npy:
            raise ValueError("fv3jedi_lm create: npx /= npy")
        if c.do_phy_trb == 0 and c.do_phy_mst == 0:
            c.do_phy = 0                                             # (:85)
        if c.do_phy == 1 and c.do_phy_mst == 1:
            raise RuntimeError("fv3jedi_lm create: moist physics (do_phy_mst = 1) is not built; set conf.do_phy_mst = 0")
        cfg = fv3lm.default_config(npx - 1, npz, dt=c.dt, ptop=c.ptop, **self._flags)
        self._h = fv3lm.FV3LM(cfg, c.ak, c.bk, emu=self._emu)
        h = self._h
        if self._metrics is not None:
            h.set_metrics(self._metrics)
            if all(k in self._metrics for k in ("a11", "a12", "a21", "a22")):
                o = cfg.ng - 1
                cd = lambda a: np.ascontiguousarray(a[:, o + 1:o + npx, o + 1:o + npx])
                h.set_c2l(*[h.scatter_c(cd(self._metrics[k])) for k in ("a11", "a12", "a21", "a22")])
                self._winds = True
        c.npx, c.npy, c.npz = npx, npy, npz
        c.hydrostatic = bool(cfg.hydrostatic)
        c.im, c.jm, c.lm = h.nxl, h.nyl, npz                          # physics grid of one sub-domain (:78-80)
        self._names = PROGNOSTIC[:8 if c.hydrostatic else 10]
        shape = (h.nsub, npz, h.nyl, h.nxl)
        self.traj = {k: np.zeros(shape) for k in self._names + ["ua", "va"]}       # allocate_traj (utils :127-160)
        self.traj["phis"] = np.zeros((h.nsub, h.nyl, h.nxl))
        self.pert = {k: np.zeros(shape) for k in self._names + ["ua", "va"]}       # allocate_pert (utils :75-100)
        self._phis_sent = False

    @property
    def handle(self):
        """the fv3lm.FV3LM behind this object (multi-rank callers attach the transport: comm_init_nccl / comm_set_callback)"""
        return self._h

    def delete(self):
        if self._h is not None:
            self._h.close() if hasattr(self._h, "close") else None
        self._h = None
        self.traj = {}; self.pert = {}
        self.conf.ak = self.conf.bk = None

    # ---- init / final (:106-145, :189-214) --------------------------------------------------------------------------------
    def init_nl(self):
        pass

    def init_tl(self):
        self._ipert_to_zero()

    def init_ad(self):
        self._ipert_to_zero()

    def final_nl(self):
        pass

    def final_tl(self):
        self._ipert_to_zero()

    def final_ad(self):
        self._ipert_to_zero()

    def _ipert_to_zero(self):                                         # (:241-252)
        for k in ("ua", "va"):
            if k in self.pert:
                self.pert[k][...] = 0.0

    # ---- helpers -----------------------------------------------------------------------------------------------------------
    def _fields(self, d):
        out = {}
        for k in self._names:
            a = d[k]
            if a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
                raise ValueError("fv3jedi_lm: field %s must be a contiguous float64 array" % k)
            out[k] = a
        return out

    def _send_traj(self, slot):
        h = self._h
        if not self._phis_sent:
            h.set_phis(self.traj["phis"]); self._phis_sent = True
        h.traj_set(slot, self._fields(self.traj))

    def _ltraj(self, slot):
        """set_ltraj (turbulence :375-533) for the trajectory held in `slot`; kept per time level when conf.saveltraj"""
        if self.conf.saveltraj and slot in self._ltraj_set:
            return
        if self.bl_driver is None:
            raise RuntimeError("fv3jedi_lm: do_phy_trb = 1 needs bl_driver (the nonlinear boundary-layer scheme of set_ltraj)")
        self._h.turb_set_ltraj(slot, self.bl_driver(self.traj))
        self._ltraj_set.add(slot)

    # ---- steps (:149-185) --------------------------------------------------------------------------------------------------
    def step_nl(self):
        c, h = self.conf, self._h
        self._send_traj(c.n)
        state = c.n
        if c.do_dyn == 1:
            h.step_nl(c.n, c.n + 1)
            state = c.n + 1
            h.traj_get(state, self._fields(self.traj))
            if self._winds:
                h.traj_get_winds(self.traj["ua"], self.traj["va"])
        if c.do_phy == 1 and c.do_phy_trb == 1:
            self._ltraj_set.discard(state)                            # the state of that slot has just changed
            self._ltraj(state)                                        # (local trajectory of the state the physics acts on)
            h.turb_step_nl(state, state)
            h.traj_get(state, self._fields(self.traj))
            self._ltraj_set.discard(state)                            # not reused: a later step_tl / step_ad sets its own from its traj

    def step_tl(self):
        c, h = self.conf, self._h
        self._ipert_to_zero()
        self._send_traj(c.n)
        h.pert_upload(self._fields(self.pert))
        if c.do_dyn == 1:
            h.step_tl_dev(c.n)
        if c.do_phy == 1 and c.do_phy_trb == 1:
            self._ltraj(c.n)
            h.turb_step_tl_dev(c.n)
        h.pert_download(self._fields(self.pert))
        self._ipert_to_zero()

    def step_ad(self):
        c, h = self.conf, self._h
        self._ipert_to_zero()
        self._send_traj(c.n)
        h.pert_upload(self._fields(self.pert))
        if c.do_phy == 1 and c.do_phy_trb == 1:
            self._ltraj(c.n)
            h.turb_step_ad_dev(c.n)
        if c.do_dyn == 1:
            h.step_ad_dev(c.n)
        h.pert_download(self._fields(self.pert))
        self._ipert_to_zero()
