import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Make sure the oracles (and, if nvcc is around, the engine) are built."""
    from gblastn_b200 import build
    build.build_oracles()
    try:
        build.build_engine()
    except Exception:      # no nvcc on this box: the prebuilt .so travels with the snapshot
        pass
    return True
