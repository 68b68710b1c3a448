"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/maz_tree.h declares; the ctypes table matches the header; without a GPU every compute entry
point fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    out = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if not fn.endswith(".h"):
            continue
        src = open(os.path.join(ROOT, "include", fn)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        out |= set(re.findall(r"\b(maz_[a-z0-9_]+)\s*\(", src))
    return out


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in sorted(syms):
        assert hasattr(lib, s), f"{s} declared in include/*.h but not exported by {built_lib}"


def test_ctypes_table_matches_header(built_lib):
    from mazero_b200 import _lib

    assert set(_lib.SIGNATURES) == header_symbols()
    assert _lib.lib.maz_abi_version() == 1


def test_no_cpu_fallback(built_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mazero_b200 import cytree

    with pytest.raises(RuntimeError, match="CUDA|device"):
        cytree.Tree_batch(4, 1, 3, 2, 5, 0.01, 0, 0.75, 0.8)


def test_product_does_not_import_oracle():
    """The product path must never route through the CPU checker."""
    pkg = os.path.join(ROOT, "mazero_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower() or f == "build.py", f"{f} mentions the oracle"


def test_bad_arguments_rejected_before_touching_cuda(built_lib):
    from mazero_b200 import _lib

    h = ctypes.c_void_p()
    assert _lib.lib.maz_tree_create_ex(ctypes.byref(h), 4, 1, 3, 64, 5, 0.01, 0, 0.75, 0.8, 0, 0) == _lib.MAZ_ERR_UNSUPPORTED
    assert "sampled_times" in _lib.last_error()
    assert _lib.lib.maz_tree_create_ex(ctypes.byref(h), 0, 1, 3, 2, 5, 0.01, 0, 0.75, 0.8, 0, 0) == _lib.MAZ_ERR_INVALID
    assert _lib.lib.maz_tree_prepare(None, None, None, None, None, 1, 0.0, None) == _lib.MAZ_ERR_INVALID
