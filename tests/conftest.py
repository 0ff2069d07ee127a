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
    from tests import oracle_util
    return oracle_util.load()


@pytest.fixture(scope="session")
def engine_lib():
    import codes_b200
    if not os.path.exists(codes_b200.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return codes_b200.load_library()
