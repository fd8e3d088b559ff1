"""ORACLE shim: synthetic grid metrics live in fv3-jedi-linearmodel_b200/synth/grid.py."""
from synth.grid import *  # noqa: F401,F403
