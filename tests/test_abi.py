"""The C-ABI boundary: libfv3lm_b200.so (nvcc, sm_100a) must load on a box without a GPU and export every
function include/fv3lm_b200.h declares; creating a handle without a device must fail loudly (no CPU fallback)."""
import ctypes
import os
import re
import pytest
import fv3lm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "fv3lm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fv3lm_[a-z0-9_]+)\s*\(", src)) - {"fv3lm_exchange_fn"})


def test_header_symbols_exported_by_product_library():
    path = fv3lm.lib_path(emu=False)
    if not os.path.exists(path):
        pytest.skip("product library not built (run __graft_entry__.build())")
    lib = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "missing export " + s
    for s in fv3lm.EXPORTS:
        assert s in syms, "python binding uses an undeclared symbol " + s


def test_header_symbols_exported_by_emulation_library():
    lib = ctypes.CDLL(fv3lm.lib_path(emu=True))
    for s in declared_symbols():
        assert hasattr(lib, s), "missing export " + s


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    path = fv3lm.lib_path(emu=False)
    if not os.path.exists(path):
        pytest.skip("product library not built")
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        fv3lm.FV3LM(fv3lm.default_config(12, 4))


def test_bad_decomposition_is_an_error():
    with pytest.raises(RuntimeError):
        fv3lm.FV3LM(fv3lm.default_config(12, 4, rank=0, nranks=5), emu=True)
    with pytest.raises(RuntimeError):
        fv3lm.FV3LM(fv3lm.default_config(12, 4, layout_x=5, layout_y=1), emu=True)


def test_unsupported_flags_are_errors():
    """the reference aborts on inconsistent set-ups (src/fv3jedi_lm_mod.F90:91-94); here create returns an error status"""
    for kw in (dict(hord_mt=9), dict(hord_tr=12), dict(nq=5), dict(n_split=0), dict(nord=4), dict(dt=0.0), dict(q_split_max=9),
               dict(q_split_dynamic=2), dict(beta=-0.2), dict(beta=1.0), dict(d_ext=-0.02)):
        with pytest.raises(RuntimeError):
            fv3lm.FV3LM(fv3lm.default_config(12, 4, **kw), emu=True)
    # beta > 0 (grad1_p_update / split_p_grad) and the reference's default d_ext = 0.02 are built
    for kw in (dict(beta=0.4), dict(d_ext=0.02), dict(d_ext=0.02, hydrostatic=0)):
        fv3lm.FV3LM(fv3lm.default_config(12, 4, **kw), emu=True)
    with pytest.raises(RuntimeError):
        fv3lm.FV3LM(fv3lm.default_config(12, 120), emu=True)
