import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(autouse=True)
def _wavefront_on_small_frames(monkeypatch):
    """Test frames are tiny: with the library's default a bounce's wavefront list below 16384 samples is finished by one
    PathTail launch (nrt_renderer.h: hardTailBelow), so the wavefront kernels production frames run would go untested.
    Tests therefore run with the short-list shortcut off; the tests OF the shortcut remove the variable again."""
    monkeypatch.setenv("NRT_HARD_TAIL_BELOW", "0")
