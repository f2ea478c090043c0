import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def evp_lib():
    """The product library; built in-tree by __graft_entry__.build().  Never silently absent."""
    from mpas_seaice_b200 import host
    if not os.path.exists(host.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return host.load_library()
