import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cuda-fft-convolution_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def fc():
    """The product package (ctypes binding over libfftconv.so)."""
    import fftconv_b200
    return fftconv_b200


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    return o
