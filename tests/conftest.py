import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure both shared libraries exist (no-op when they are up to date)."""
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle
    return oracle.load()


def have_reference() -> bool:
    return os.path.isdir(os.path.join(REFERENCE, "scene"))


needs_reference = pytest.mark.skipif(not have_reference(), reason="reference tree not mounted (GPU box)")
