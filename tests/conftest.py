import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    # the torch references must be real fp32 (TF32 would be less accurate than the kernels under test)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from countr_b200 import _lib
    _lib.require_device()  # raises if the .so is missing or the device is not sm_100
    return torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _deterministic_draws(request):
    """Every test draws its random tensors from a seed derived from its own id: a threshold that holds for one draw and not
    for another is a test bug, and the suite must give the same verdict on every fresh box."""
    import zlib
    import torch
    seed = zlib.crc32(request.node.nodeid.encode()) & 0x7FFFFFFF
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    yield
