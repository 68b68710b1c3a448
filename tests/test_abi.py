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


def test_turn_infer_and_hostrng_entry_points_validate_arguments(built_lib):
    """Argument checks of include/maz_turn.h, maz_infer.h and maz_hostrng.h happen before any CUDA call: they can be
    exercised without a GPU (reference convention: invalid input -> error code + message, never a crash)."""
    from mazero_b200 import _lib

    lib = _lib.lib
    assert lib.maz_root_prepare_dev(None, None, None, 4, 2, 3, -1, 0.25, 1.0, None, None, None, None, None) == _lib.MAZ_ERR_INVALID
    x = np.zeros(4096, np.float32)
    p = ctypes.c_void_p(x.ctypes.data)
    assert lib.maz_root_prepare_dev(p, None, p, 4, 2, 300, -1, 0.25, 1.0, p, p, p, None, None) == _lib.MAZ_ERR_UNSUPPORTED
    assert "256" in _lib.last_error()
    assert lib.maz_root_prepare_dev(p, None, p, 4, 2, 3, 2, 0.25, 1.0, p, p, p, None, None) == _lib.MAZ_ERR_INVALID   # cur >= agents
    assert lib.maz_agent_turn_dev(7, 4, 2, 3, 5, 0, p, p, p, p, None, 1.0, None, 0.0, None, None, p, p, p, None, None) == _lib.MAZ_ERR_INVALID
    assert lib.maz_agent_turn_dev(1, 4, 2, 3, 5, 0, p, p, p, p, None, 1.0, None, 0.0, None, None, p, p, p, None, None) == _lib.MAZ_ERR_INVALID
    assert "uniforms" in _lib.last_error()
    assert lib.maz_dirichlet_dev(p, 4, 300, 0.3, 1, None) == _lib.MAZ_ERR_INVALID
    assert lib.maz_mlp_forward(None, p, 4, p, None) == _lib.MAZ_ERR_INVALID
    assert lib.maz_mlp_recurrent(None, None) == _lib.MAZ_ERR_INVALID
    assert lib.maz_infer_recurrent(None, None) == _lib.MAZ_ERR_INVALID
    assert lib.maz_infer_recurrent_small(None, None) == _lib.MAZ_ERR_INVALID
    assert lib.maz_infer_small_nq() in (4, 8)
    key = np.zeros(624, np.uint32)
    pos = ctypes.c_int(624)
    out = np.zeros(64, np.float32)
    kp, op = ctypes.c_void_p(key.ctypes.data), ctypes.c_void_p(out.ctypes.data)
    assert lib.maz_legacy_dirichlet(kp, ctypes.byref(pos), 1.5, 8, 8, op, None, 1) == _lib.MAZ_ERR_UNSUPPORTED   # numpy's other branch
    assert lib.maz_legacy_dirichlet(kp, ctypes.byref(pos), 0.3, 0, 8, op, None, 1) == _lib.MAZ_ERR_INVALID
    pos = ctypes.c_int(700)
    assert lib.maz_legacy_dirichlet(kp, ctypes.byref(pos), 0.3, 8, 8, op, None, 1) == _lib.MAZ_ERR_INVALID
