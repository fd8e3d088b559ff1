"""ORACLE shim: the python cubed-sphere index maps (restatement of the FMS halo semantics,
tools/fv_mp_nlm_mod.F90:285-1471) live in fv3-jedi-linearmodel_b200/synth/cubed_sphere.py because the
synthetic grid generator needs them too; they are never used by the CUDA path."""
from synth.cubed_sphere import *  # noqa: F401,F403
from synth.cubed_sphere import _assign, _side_of  # noqa: F401
