import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """libqcknot.so must exist before any test touches the package (nvcc cross-compiles without a GPU)."""
    import qcknot

    qcknot.build_library()
    return qcknot.LIB_PATH
