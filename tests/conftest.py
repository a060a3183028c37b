"""pytest configuration: the `gpu` marker and import paths.

`-m "not gpu"`: oracle vs golden vectors, host logic, C-ABI symbol checks (no CUDA device needed).
`-m gpu`      : parity tests proper -- the CUDA path, called through the C ABI, against the oracle.
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "python-wlsqm_b200", ROOT / "oracle"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

os.environ.setdefault("OMP_WAIT_POLICY", "passive")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        from wlsqm_b200 import _lib
        return _lib.lib().wlsqm_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture
def rng():
    """seed 42, like the reference's tests (tests/conftest.py:17-23)"""
    return np.random.default_rng(42)
