import os
import sys

import pytest

# The pre-existing suites assert BITWISE equality with the oracle: they run the library in its bitwise
# arithmetic ("exact" = 1).  The tolerance mode (the library default, k_march) has its own suites
# (test_fast_formulas.py, test_gpu_fast.py, the multi-GPU cases of run_multi_gpu.py) that pass exact=False.
os.environ.setdefault("SWCU_EXACT", "1")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def swlib():
    """The product library; built in-tree if it is stale (nvcc cross-compiles without a GPU)."""
    from ocean_model_arch_b200 import build
    build.build()
    from ocean_model_arch_b200 import _lib
    return _lib.lib()


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.init()
    return torch.device("cuda", 0)
