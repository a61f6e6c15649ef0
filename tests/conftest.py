import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def lib():
    """The built C-ABI library (built on demand with nvcc; cross-compiles without a GPU)."""
    from flash_attention_from_scratch_b200 import _lib, build

    build.build()
    return _lib.load()


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["cfg1_bf16_1x128x2", "bf16_2x256x3", "fp16_2x256x3", "bf16_1x512x1", "bf16_1x1280x1"]


@pytest.fixture(scope="session", params=GOLDEN_CASES)
def golden(request):
    import torch

    return torch.load(os.path.join(GOLDEN_DIR, request.param + ".pt"), weights_only=False)
