"""pytest configuration: the ``gpu`` marker, oracle / golden fixtures."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    from oracle import Orc
    return Orc()


@pytest.fixture(scope="session")
def ref():
    from oracle import Ref
    if not Ref.available():
        pytest.skip("oracle/_ref/libfssref.so not built (needs /root/reference)")
    return Ref()


@pytest.fixture(scope="session")
def golden():
    from golden_util import Golden
    return Golden()
