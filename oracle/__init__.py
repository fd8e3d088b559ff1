"""ORACLE: CPU restatement of the reference algorithm. Test infrastructure only --
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs, never by the product path (fv3-jedi-linearmodel_b200/).

PARITY UNPINNED: the reference (l90lpa/fv3-jedi-linearmodel) ships no tests, golden vectors
or fixtures, and it cannot be compiled in this image (no Fortran compiler, FMS, MPI, NetCDF;
SURVEY.md section 8c), so these restatements follow the hand-written nonlinear sources
(model/*_nlm.F90, file:line cited per function) and are checked by self-consistency only:
dot-product (adjoint) tests, Taylor tests against finite differences, conservation /
symmetry invariants and decomposition invariance.  TL = torch.func.jvp and AD =
torch.func.vjp of the nonlinear restatement (exact derivatives, split_* = .false.)."""
