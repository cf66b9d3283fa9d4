import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure only), built on demand with g++."""
    from oracle import oracle_py
    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def cuda_lib():
    """libcpmpc.so, built on demand with nvcc (cross-compiles without a GPU)."""
    from cart_pole_mpc_b200 import build, _capi
    build.build_library()
    return _capi.lib()


def load_golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name))
