import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running GPU sweep")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o

    o.build()
    return o


@pytest.fixture(scope="session")
def dct():
    """The product package; GPU tests call through its C ABI."""
    import torch

    import cuda_dct_idct_b200 as m

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    m.lib()  # raises loudly if the extension is missing
    return m
