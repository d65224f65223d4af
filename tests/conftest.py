import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Both shared objects are built in-tree once per session (no-op when up to date)."""
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def golden():
    return GOLDEN


@pytest.fixture(scope="session")
def tiny(golden):
    from parafem_b200 import host
    return host.read_deck_p121(os.path.join(golden, "xx3-tiny"))


@pytest.fixture(scope="session")
def demo():
    from parafem_b200 import host
    return host.cube_p121(20, 20, 20, 20, aa=.5, bb=.5, cc=.5, round_mode=1)
