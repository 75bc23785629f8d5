"""The C-ABI product library loads on a CPU-only box and exports every symbol include/dmnd_b200.h declares; it refuses
to create a context without a CUDA device (no CPU fallback).  CPU only, no compute calls."""
import ctypes, os, re
import pytest
from conftest import ROOT, PRODUCT_LIB


def header_functions():
    txt = open(os.path.join(ROOT, "include", "dmnd_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dmnd_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    from diamond_b200 import api
    assert header_functions() == sorted(api.SYMBOLS)


def test_product_library_exports_every_symbol(product_lib):
    for name in header_functions():
        assert hasattr(product_lib, name), name
    assert product_lib.dmnd_backend().decode() == "cuda-sm100a"


def test_product_library_has_no_cpu_path(product_lib):
    from diamond_b200 import api
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    with pytest.raises(api.DmndError, match="no CUDA device|CUDA"):
        api.Context(product_lib)


def test_product_library_does_not_contain_the_oracle(product_lib):
    # the oracle identifies itself through dmnd_backend(); the product must not carry its translation unit
    data = open(PRODUCT_LIB, "rb").read()
    assert b"oracle-cpu" not in data


def test_struct_layouts():
    from diamond_b200 import api
    assert ctypes.sizeof(api.Hit) == 16 and api.HIT_DTYPE.itemsize == 16
    assert ctypes.sizeof(api.DpProblem) == 16 and api.PROBLEM_DTYPE.itemsize == 16
    assert ctypes.sizeof(api.DpResult) == api.RESULT_DTYPE.itemsize == 56
    assert ctypes.sizeof(api.Match) == api.MATCH_DTYPE.itemsize
