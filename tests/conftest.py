import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session")
def engine():
    """The C ABI (libopflow_b200.so) initialised on cuda:0 -- GPU tests only."""
    from opflow_b200 import capi
    l = capi.lib()
    capi.check(l.opf_init(0))
    yield l


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O
