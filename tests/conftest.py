import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the native pieces exist (no-op when already built; the .so files travel to the GPU box)."""
    import __graft_entry__ as ge
    if not (os.path.isfile(ge.LIB) and os.path.isfile(ge.HOST_LIB)):
        ge.build()
    yield


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "splat_slice_golden.npz"))
