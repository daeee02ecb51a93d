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
    """The oracle (gcc) is always built; the CUDA library is built when nvcc is here, and must exist on a GPU box."""
    from oracle import pyoracle
    pyoracle.build()
    from genome_b200 import build as gb_build
    try:
        gb_build.build()
    except RuntimeError:
        if not os.path.exists(gb_build.LIB):
            raise
    yield


@pytest.fixture(scope="session")
def gpu():
    import ctypes as C
    from genome_b200 import capi
    n = C.c_int(0)
    capi.check(capi.lib().gb_device_count(C.byref(n)))
    assert n.value >= 1, "no CUDA device: the product has no CPU fallback"
    return n.value
