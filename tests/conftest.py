import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def xg_lib():
    from xcltk_b200 import lib
    return lib.load()


@pytest.fixture(scope="session")
def gpu_ctx():
    from xcltk_b200 import engine
    return engine.get_context(0)
