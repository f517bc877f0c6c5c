import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def cuda_lib():
    """The in-tree CUDA library; GPU tests fail (not skip) if it is missing."""
    import torch
    from opfgym_b200 import capi
    assert torch.cuda.is_available(), "GPU test tier needs a CUDA device"
    return capi.load()
