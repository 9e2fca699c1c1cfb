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
    import gadm_b200  # noqa: F401
    from gadm_b200 import _lib
    _lib.ensure_init(0)          # raises loudly if libgadm.so is missing or the GPU is not sm_100
    return torch.device("cuda", 0)
