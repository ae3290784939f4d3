import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The compiled reference (oracle/_ref).  Present wherever it was prebuilt."""
    from oracle.oracle import Reference
    if not Reference.available():
        try:
            from oracle.oracle import build
            build()
        except Exception:
            pass
    if not Reference.available():
        pytest.skip("oracle/_ref not built (needs the reference tree at build time)")
    return Reference()
