"""Synthetic-input generation for tests and bench.py: cubed-sphere grid metrics (what the
host model's gridstruct would supply through fv3lm_set_metric), python halo index maps used
to fill metric halos, and synthetic model states.  Not on the compute path: the library's own
halo maps live in csrc/mosaic.cu."""
