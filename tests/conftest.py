import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The fp64 CPU restatement (test infrastructure; built on demand)."""
    from oracle import oracle as o
    o.lib()
    return o


@pytest.fixture(scope="session")
def mg():
    import mgodpl_b200
    return mgodpl_b200


@pytest.fixture(scope="session")
def gpu(mg):
    """The product library on a CUDA device; GPU tests fail loudly if the extension is missing."""
    mg.capi.lib()
    if mg.capi.lib().mgodpl_device_count() <= 0:
        pytest.fail("no CUDA device visible: -m gpu tests must run on the GPU box")
    return mg
