import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, 'tests')):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def cuda_lib():
    """Loads the in-tree C-ABI library; GPU tests must fail loudly (not skip) when it is missing."""
    import torch
    import tamtr_b200
    assert torch.cuda.is_available(), "GPU test selected but no CUDA device is visible"
    tamtr_b200._lib.lib()
    # fp32 parity is checked against true-fp32 library GEMMs/convs (the reference's CPU path has no TF32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return tamtr_b200
