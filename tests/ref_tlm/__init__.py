"""TEST INFRASTRUCTURE ONLY: statement-by-statement numpy transliterations of Tapenade tangent-mode routines of the reference
(model_tlmadm/*_tlm.F90).  They pin the ORACLE (oracle/, whose TL is torch.func.jvp of a restated primal) to the reference's own
tangent code: which variables are active, which branch is differentiated, the cube-edge special cases.  Nothing under
fv3-jedi-linearmodel_b200/ imports this package.  The reference cannot be compiled in this image (no Fortran compiler, FMS, MPI), so
these and the reference-executed
fixtures of tests/golden/ref_*.npz (transpiler f90py.py in this package, generation time only) are the reference-derived pins."""
import numpy as np


class F:
    """Fortran-style array with arbitrary lower bounds: F((lo, hi), (lo, hi), ...); a[i, j] uses Fortran indices."""

    def __init__(self, *bounds, data=None):
        self.lo = tuple(b[0] for b in bounds)
        shape = tuple(b[1] - b[0] + 1 for b in bounds)
        self.a = np.zeros(shape) if data is None else np.asarray(data, dtype=np.float64).reshape(shape)

    def _ix(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        out = tuple(i - l for i, l in zip(idx, self.lo))
        assert all(o >= 0 for o in out), (idx, self.lo)
        return out

    def __getitem__(self, idx):
        return self.a[self._ix(idx)]

    def __setitem__(self, idx, v):
        self.a[self._ix(idx)] = v

    def fill(self, v):
        self.a[...] = v
