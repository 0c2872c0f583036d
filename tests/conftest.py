import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 0x1CE27002025


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def engine_lib():
    """The CUDA engine, built in-tree.  GPU tests must exercise it; there is no fallback."""
    from incerto_b200 import _build, _ffi
    _build.build()
    return _ffi.load()


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle
    return oracle.load()
