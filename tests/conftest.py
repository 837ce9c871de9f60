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
    from oracle import bindings as ob
    return ob.Oracle()


@pytest.fixture(scope="session")
def reference():
    """the unmodified reference compiled from source (oracle/_ref); skipped where it was never built"""
    from oracle import bindings as ob
    if not ob.Reference.available():
        pytest.skip("oracle/_ref/librayforce_ref.so not built (reference sources not mounted)")
    return ob.Reference.get()


@pytest.fixture(scope="session")
def ctx():
    """one rfb context on cuda:0 for the whole GPU session; fails loudly when the library or the device is missing"""
    import torch
    from rayforce_b200 import Context
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    c = Context(0)
    yield c
    c.close()
