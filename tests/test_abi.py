"""CPU test: libaefft.so loads and exports every symbol include/aefft.h declares; compute entry points fail loudly
(AEFFT_ERR_CUDA) instead of falling back when no GPU is usable."""
import os
import re

import pytest

import aefft_ctypes as A
from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "aefft.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(aefft_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    lib = A.lib()
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/aefft.h but not exported: {missing}"
    assert lib.aefft_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-GPU behaviour is checked in the CPU tier")
    with pytest.raises(A.AefftError) as e:
        A.Ctx(0)
    assert e.value.code == A.ERR_CUDA
