import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "taiwan-tongues-asr-ce_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    # exact fp32 references: no TF32 in the torch control computations
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch.device("cuda", 0)


@pytest.fixture(scope="session")
def built_lib():
    """Path of libttasr_b200.so, building it if the tree has none (CPU boxes cross-compile with nvcc)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("ttasr_build", os.path.join(PKG, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if os.path.exists(mod.LIB) and not os.path.exists("/usr/local/cuda/bin/nvcc"):
        return mod.LIB
    return mod.build(verbose=False)
