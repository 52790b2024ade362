import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/), compiled on first use."""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def cabi():
    """ctypes binding of the product's C-ABI (host entry points work without a GPU)."""
    from klujax_b200 import build, _cabi
    if not os.path.exists(_cabi.LIB_PATH):
        build.build_cuda()
    return _cabi
