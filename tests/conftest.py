import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_device() -> bool:
    try:
        from fringe_b200 import engine
        return engine.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not errored) on a machine without a CUDA device."""
    if _have_device():
        return
    skip = pytest.mark.skip(reason="no CUDA device on this machine (fringe_b200 has no CPU path)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_report_header(config):
    import oracle
    kind = "reference-header build (oracle/_ref)" if oracle.available("reference") else \
        "PORT ONLY -- oracle/_ref/libfringe_ref.so is absent, parity is checked against the restatement alone"
    return f"oracle: {kind}"


@pytest.fixture(scope="session")
def oracle_lib():
    """CPU oracle (test infrastructure): the reference-header build when present, else the port -- and then
    it says so, because the pin tests would otherwise compare the port with itself."""
    import warnings

    import oracle
    if not oracle.available("reference"):
        warnings.warn("oracle/_ref/libfringe_ref.so is absent: parity tests run against the restated port only")
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
