import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib_built():
    """Make sure libb200spec.so exists (cross-compiles without a GPU)."""
    from audio_tabs_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def cuda_device(lib_built):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible (no CPU fallback exists)")
    return torch.device("cuda", 0)
