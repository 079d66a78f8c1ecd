import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built_lib():
    """Path of the in-tree C-ABI library; builds it if missing (nvcc cross-compiles
    sm_100a without a GPU)."""
    from zfista_b200 import build

    return build.build()


@pytest.fixture(scope="session")
def gpu(built_lib):
    """The CUDA path must be the one that runs: fail (not skip) when the library or
    the device is missing on a box that is supposed to have one."""
    from zfista_b200 import _lib

    n = _lib.lib().zf_device_count()
    assert n >= 1, "no CUDA device visible: GPU tests cannot run (there is no CPU fallback)"
    return n
