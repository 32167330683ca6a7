import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def built_libraries():
    """Build the host library, the product library (nvcc cross-compiles without a GPU) and the oracle."""
    import subprocess

    from longcallr_b200 import build

    build.build_all()
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    yield
