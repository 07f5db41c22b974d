"""pytest configuration: `gpu` marker, import paths, a shared Device fixture for the -m gpu tests."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def dev():
    """One context/stream for the whole session.  Fails loudly (no CPU fallback) if the extension or a GPU is missing."""
    import rust_autograd_b200 as agb
    d = agb.Device(0)
    yield d
    d.close()
