"""pytest configuration: `-m gpu` tests need a B200 (run under gpurun); everything else runs on CPU."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "autoencoder-fft_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (sm_100a); run with -m gpu on the B200 box")


@pytest.fixture(scope="session")
def ctx():
    """One aefft context on cuda:0.  No fallback: a missing library or GPU is a test failure, not a skip."""
    import aefft_ctypes as A

    c = A.Ctx(0)
    assert c.precision == A.PRECISION_BF16X3, "library default is the tensor-core BF16X3 mode"
    c.set_precision(A.PRECISION_FP32)  # the fp32 CUDA-core kernels are what test_coord_gpu pins; test_tc_gpu switches modes
    yield c
    c.close()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference (oracle/_ref/libref.so) when it was built; None otherwise."""
    import ref_lib

    return ref_lib if ref_lib.available() else None
