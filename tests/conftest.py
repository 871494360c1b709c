import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle_built():
    """The C restatement is (re)built on demand; _ref only where /root/reference exists."""
    from oracle import bindings
    bindings.build()
    return bindings


@pytest.fixture(scope="session")
def gpu_device():
    import openmmgridforce_b200 as gf
    dev = gf.Device(0)          # raises (loudly) when the CUDA library or the GPU is missing
    yield dev
    dev.close()
