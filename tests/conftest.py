import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _pkg():
    import importlib
    return importlib.import_module("x265-yuuki-asuna_b200")


@pytest.fixture(scope="session")
def pkg():
    return _pkg()


@pytest.fixture(scope="session")
def ctx():
    """A live backend context.  GPU tests FAIL (not skip) when the CUDA library cannot start."""
    p = _pkg()
    c = p.Ctx(0)
    yield c
    c.close()
