import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (sm_100a); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def default_weights():
    import styletts_zs_b200 as stz
    return stz.init_weights(stz.DEFAULT, 0)


@pytest.fixture(scope="session")
def tiny_weights():
    import styletts_zs_b200 as stz
    return stz.init_weights(stz.TINY, 0)
