import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "python"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def plf():
    import plf as m
    return m


@pytest.fixture(scope="session")
def oracle(plf):
    # TEST INFRASTRUCTURE: build the CPU oracle on demand (g++ only)
    if not os.path.exists(plf.ORACLE_LIB):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    return plf.load_oracle()


@pytest.fixture(scope="session")
def product(plf):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return plf.load_product()   # raises if the CUDA library is not built: no fallback


@pytest.fixture(scope="session")
def pair1(plf):
    return plf.synth_pair(752, 480, 1)
