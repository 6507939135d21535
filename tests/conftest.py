import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.binding import oracle_backend
    return oracle_backend()


@pytest.fixture(scope="session")
def cuda():
    """The product backend; GPU tests fail loudly if the extension is missing."""
    from i3rc_monte_carlo_model_b200._lib import backend
    be = backend()
    assert be.device_count() > 0, "no CUDA device visible"
    return be


@pytest.fixture
def cuda_absent():
    """Skips the test on a box that has a GPU (it checks the behaviour without one)."""
    from i3rc_monte_carlo_model_b200._lib import backend
    if backend().device_count() > 0:
        pytest.skip("a CUDA device is present")
