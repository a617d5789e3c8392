import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "scrna-parameter-estimation_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture
def tuning(monkeypatch):
    """``tuning(MM_MOMENTS_KERNEL="tile", ...)``: set MM_* variables and make the library re-read them (they are
    read once at load); the library goes back to the plain environment when the test ends."""
    from memento_b200 import _lib

    def apply(**env):
        for k, v in env.items():
            monkeypatch.setenv(k, str(v))
        _lib.reload_tuning()

    yield apply
    monkeypatch.undo()
    _lib.reload_tuning()
