import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("this test is marked gpu and needs a CUDA device")
    from bevfusion_b200 import _C
    _C.lib()  # the CUDA library must load: no fallback
    return torch.device("cuda:0")


def ref_module(name):
    """Reference extension from oracle/_ref (None if it was not built)."""
    from oracle.build_ref import built, load_ref
    return load_ref(name) if built(name) else None
