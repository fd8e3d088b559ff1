"""ORACLE: CPU restatement of the reference algorithm. Test infrastructure only --
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs, never by the product path (fv3-jedi-linearmodel_b200/).

PARITY PINNED to the reference's own code, executed: the reference (l90lpa/fv3-jedi-linearmodel) ships
no tests or golden vectors and cannot be compiled in this image (no Fortran compiler, FMS, MPI, NetCDF;
SURVEY.md section 8c), so tests/ref_tlm/f90py.py transpiles its Fortran (the Tapenade-generated
tangent-linear AND reverse-mode sources) to Python and tests/golden/make_ref_golden.py runs it in the
build container on seeded inputs; the outputs are the fixtures tests/golden/ref_*.npz, which these
restatements reproduce (tests/test_ref_golden.py):
  D_SW_TLM, D_SW_FWD/BWD                       values bit for bit, tangents / adjoints <= 6e-15 (five configurations)
  DYN_CORE_TLM (six tiles)                     <= 2e-13 / 5e-15: non-hydrostatic, hydrostatic, beta = 0.4, d_ext = 0.02 (six configurations)
  DYN_CORE_FWD/BWD (reverse sweep, six tiles)  all six adjoints <= 2e-15 once the w-damping adjoint is left out as dyn_core_adm.F90 does (reference quirk, see the test)
  FV_DYNAMICS_TLM (whole step)                 <= 6e-13 / 7e-15, non-hydrostatic and hydrostatic
  FV_DYNAMICS_FWD/BWD (whole hydrostatic step) adjoints of all eight prognostics 3e-15
plus 22 routines transliterated by hand (tests/ref_tlm/*.py, tests/test_ref_tlm.py).  Not pinned by
reference code: c2l_ord4 and the physics (turbulence) -- their pins are named in their own headers.
The restatements follow the hand-written nonlinear sources (model/*_nlm.F90, file:line cited per
function); TL = torch.func.jvp and AD = torch.func.vjp of them (two-sided mode: a Splice autograd
function gives "value of the trajectory scheme, derivative of the perturbation scheme")."""
