import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """libmaz_b200.so, (re)built with nvcc when stale."""
    from mazero_b200 import build

    return build.build()


@pytest.fixture(scope="session")
def oracle_built():
    from oracle import pyoracle

    pyoracle.build()
    return pyoracle
