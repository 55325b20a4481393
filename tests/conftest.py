import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """Build (or reuse) the in-tree native libraries once per session."""
    from opencl_dpm_b200 import build

    build.build_all()
    from oracle import oracle

    oracle.build()
    yield


def has_gpu() -> bool:
    try:
        from opencl_dpm_b200 import capi

        return capi.device_count() > 0
    except Exception:
        return False
