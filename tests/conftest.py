import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_lib():
    """CPU oracle (test infrastructure): the reference-header build when present, else the port."""
    import oracle
    return oracle.load()


@pytest.fixture(scope="session")
def ctx():
    from fringe_b200.engine import Context
    c = Context(0)
    yield c
    c.close()


def wrapped_diff(a, b):
    """|angle(a * conj(b))| for complex arrays."""
    return np.abs(np.angle(a * np.conj(b)))
