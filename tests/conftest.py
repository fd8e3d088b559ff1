import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "fv3-jedi-linearmodel_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def pytest_sessionfinish(session, exitstatus):
    """table of the achieved parity errors (tests/common.py: record) -> $FV3LM_PARITY_OUT"""
    out = os.environ.get("FV3LM_PARITY_OUT")
    if not out:
        return
    try:
        import json
        import common
        if common.RECORD:
            os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
            with open(out, "w") as fh:
                json.dump({"device": "cuda" if _has_gpu() else "host emulation", "errors": common.RECORD}, fh, indent=1, sort_keys=True)
    except Exception as e:  # noqa: BLE001
        print("parity table not written:", e)
