import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def cuda_lib():
    """The C-ABI library on a CUDA device; GPU tests fail loudly (no skip) when either is missing."""
    import torch
    from pixparse_b200 import _lib
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    lib = _lib.lib()
    _lib.check(lib.b200_device_check(), "b200_device_check")
    return lib
