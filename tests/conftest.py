import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the checker (gcc) and make sure the CUDA library exists (it is built by __graft_entry__.build())."""
    from oracle import pyoracle
    pyoracle.build()
    lib = os.path.join(ROOT, "orz_b200", "liborz_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__ as g
        g.build()
    yield
