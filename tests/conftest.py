import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import p2o
    p2o.load()
    return p2o


@pytest.fixture(scope="session")
def ctx():
    import plonky2_goldibear_b200 as p2g
    c = p2g.Context(0)  # raises loudly without the built library or without a GPU
    yield c
    c.close()
