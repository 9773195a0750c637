import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


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
    return tamtr_b200
