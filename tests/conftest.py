import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ref_available():
    from oracle import refbind
    return refbind.available("plain")


@pytest.fixture(scope="session")
def ref():
    """oracle/_ref - the unmodified reference compiled as the CPU oracle."""
    from oracle import refbind
    if not refbind.available("plain"):
        pytest.skip("oracle/_ref not built (needs /root/reference: make -C oracle ref)")
    return refbind
