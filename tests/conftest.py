import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 via gpurun)")


@pytest.fixture(scope="session")
def mmslib():
    """The built product library; GPU tests fail loudly when it is missing."""
    from mosmicrosim_b200 import backend
    return backend.lib()


@pytest.fixture(scope="session")
def gpu(mmslib):
    from mosmicrosim_b200 import backend
    n = backend.device_count()
    if n < 1:
        pytest.fail("no CUDA device visible: the -m gpu tests must run on a GPU box")
    return n
