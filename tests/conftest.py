import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the reference tree at /root/reference (build container only)")


@pytest.fixture(scope="session")
def reference_pkg():
    """The real reference (Python package + oracle/_ref build); skips where it cannot exist."""
    from oracle import reference

    mod = reference.load()
    if mod is None:
        pytest.skip("reference tree / oracle/_ref build not available")
    return mod
