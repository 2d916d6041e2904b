import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def po():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def api():
    from iris_lama_b200 import api as a
    return a


@pytest.fixture(scope="session")
def gpu_api(api):
    if api.device_count() < 1:
        pytest.fail("a -m gpu test ran without a CUDA device (the lama_b200 hot path has no CPU fallback)")
    return api


@pytest.fixture(scope="session")
def synth():
    from iris_lama_b200 import synth as s
    return s
