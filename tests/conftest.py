import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import __graft_entry__ as ge  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pt():
    if not os.path.exists(os.path.join(ge.PKG_DIR, "lib", "libptb200.so")) or not os.path.exists(os.path.join(ge.PKG_DIR, "lib", "libptb200_host.so")):
        ge.build()
    return ge.load_package()


@pytest.fixture(scope="session")
def orc():
    return ge.load_oracle()


@pytest.fixture(scope="session")
def ctx(pt):
    c = pt.Context(0)
    yield c
    c.close()
