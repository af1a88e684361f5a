import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU and the built libisdf_b200.so")


@pytest.fixture(scope="session")
def ops():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import fft_isdf_scratch_b200.kernels as K
    return K.IsdfOps(0)
