import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    from tests.helpers import load_golden
    return load_golden()


@pytest.fixture(scope="session")
def methane():
    """Methane / 3-21G AO integrals of the reference fixtures (oracle integral code)."""
    from tests.helpers import methane_integrals
    return methane_integrals()
