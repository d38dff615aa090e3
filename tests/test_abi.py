"""C-ABI presence: every symbol declared in include/*.h is exported by the built libraries (no compute calls)."""
import ctypes
import importlib
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rdp = importlib.import_module("parallel-rdp_b200")


def _declared(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"\b(%s[a-z_0-9]+)\s*\(" % prefix, text)))


def _product_lib():
    if not os.path.exists(rdp.PRODUCT_LIBRARY):
        import __graft_entry__ as g
        g.build_product()
    return ctypes.CDLL(rdp.PRODUCT_LIBRARY)


def test_product_library_exports_public_abi():
    lib = _product_lib()
    names = _declared("rdp_b200.h", "rdp_b200_")
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"librdp_b200.so does not export {n}"


def test_product_library_exports_backend_abi():
    lib = _product_lib()
    names = _declared("rdp_b200_backend.h", "rdpb_")
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"librdp_b200.so does not export {n}"
    assert lib.rdpb_backend_name is not None
    lib.rdpb_backend_name.restype = ctypes.c_char_p
    assert lib.rdpb_backend_name() == b"cuda-sm_100a"


def test_product_fails_loudly_without_cuda():
    """No silent CPU fallback: without a CUDA device the processor reports unsupported and the binding raises."""
    import numpy as np
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    _product_lib()
    with pytest.raises(RuntimeError):
        rdp.CommandProcessor(np.zeros(4 << 20, dtype=np.uint8), 2 << 20)


def test_missing_library_raises(tmp_path):
    with pytest.raises(RuntimeError):
        rdp.load_library(str(tmp_path / "nope.so"))


def test_hostsim_exports_same_abi():
    from sidebyside import build_hostsim
    lib = ctypes.CDLL(build_hostsim())
    for n in _declared("rdp_b200.h", "rdp_b200_") + _declared("rdp_b200_backend.h", "rdpb_"):
        assert hasattr(lib, n), f"hostsim does not export {n}"
