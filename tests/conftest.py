import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        from phantomsdr_b200 import _ffi

        return _ffi.LIB_PATH.exists() and _ffi.lib().b200_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_required():
    """GPU tests must FAIL, not skip, when the CUDA engine is missing on a GPU box."""
    from phantomsdr_b200 import _ffi

    assert _ffi.LIB_PATH.exists(), f"{_ffi.LIB_PATH} missing - run __graft_entry__.build()"
    n = _ffi.lib().b200_device_count()
    assert n > 0, "no CUDA device visible"
    return n
