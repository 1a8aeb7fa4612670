import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "refcheck: needs /root/reference (build container only)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import c_oracle
    c_oracle.build()
    return c_oracle


@pytest.fixture(scope="session")
def cuda_backend():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from popnet_b200._cuda_backend import CudaBackend
    return CudaBackend()
