import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """One device context for the whole GPU session (1 thread : 1 device : 1 stream)."""
    import taper_b200
    c = taper_b200.Ctx(0)
    yield c
    c.close()
