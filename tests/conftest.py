import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Compile the CUDA library and the oracle if they are missing or stale (both build without a GPU)."""
    entry.build_cuda()
    entry.build_oracle()
    return True


@pytest.fixture(scope="session")
def oracle_lib(built):
    from mpimc_b200 import lib
    return lib.ImcLib(entry.ORACLE_LIB)


@pytest.fixture(scope="session")
def gpu_lib(built):
    """The product library.  Fails (does not skip, does not fall back) when it cannot be loaded."""
    from mpimc_b200 import lib
    return lib.ImcLib(entry.LIB)
