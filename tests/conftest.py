import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: CPU test that takes > 30 s")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def cpu_kernels(monkeypatch):
    """Install tests/cpu_backend.py over valor_b200.kernels (host-logic tests only)."""
    import types
    from tests import cpu_backend
    import valor_b200.kernels as K
    for name in dir(cpu_backend):
        obj = getattr(cpu_backend, name)
        if isinstance(obj, types.FunctionType) and not name.startswith("_"):
            monkeypatch.setattr(K, name, obj)
    return K
