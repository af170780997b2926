import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def gpu():
    """One GpuInstance per session (the reference's tests are #[serial] on one adapter)."""
    import wgmath_b200 as w
    return w.GpuInstance.new()


@pytest.fixture(scope="session")
def shapes():
    import wgmath_b200 as w
    return w.ViewShapeBuffers.new()
