import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.binding import Oracle, build

    build()
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle.binding import Ref

    r = Ref()
    if not r.available:
        pytest.skip("oracle/_ref/libg4s_ref.so not built (reference tree absent)")
    return r
