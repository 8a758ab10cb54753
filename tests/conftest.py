import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long CPU test (still part of the default CPU suite unless deselected)")


def pytest_collection_modifyitems(config, items):
    """plain `pytest tests` on a box without a GPU: gpu-marked tests are skipped instead of failing in grail_cuda_create
    (the library has no CPU path, by design)"""
    try:
        import grail_rs_b200 as g
        n = g._ffi.lib().grail_cuda_device_count()
    except Exception:
        n = 0
    if n > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (the product has no CPU path)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def voice(oracle):
    return oracle.generic_voice()
