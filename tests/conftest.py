import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure), built on demand."""
    from oracle import oracle as orc

    orc.build()
    return orc


@pytest.fixture(scope="session")
def gsb_lib():
    """The product shared library, built on demand (nvcc cross-compiles without a GPU)."""
    from gs2mesh_b200 import build

    build.build()
    from gs2mesh_b200 import _lib

    return _lib.lib()


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)
