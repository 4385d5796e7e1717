import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib():
    from lara_b200 import _lib
    return _lib.load()


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference build (oracle/_ref); tests that need it skip if it is absent."""
    from oracle import ref as REF
    if not REF.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return REF.load()
