import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def emu_lib():
    """Host-emulation build of the product kernels (tests/hostemu): CPU-only check of index arithmetic and
    solver orchestration.  Never used by the product."""
    from tests.hostemu.build_hostemu import build
    from spock_b200 import _lib

    return _lib.bind(build())


@pytest.fixture(scope="session")
def cuda_lib():
    """The real CUDA library, built in-tree.  GPU tests fail loudly if it is missing."""
    from spock_b200 import _lib, build

    build.build()
    return _lib.load()
