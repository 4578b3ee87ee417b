import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): compiled on first use."""
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def mswb():
    """The product's C ABI through ctypes; the library must already be built (no fallback)."""
    import msweep_b200
    msweep_b200.lib()
    return msweep_b200


@pytest.fixture(scope="session")
def ctx(mswb):
    c = mswb.Context(0)
    yield c
    c.close()
