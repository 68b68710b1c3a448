"""GPU: the fused tensor-core inference path against a plain PyTorch fp32 reference."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def pack_operand(w):
    """[rows, K] -> bf16 in the kernels' K-major core-matrix layout (mazero_b200/csrc/umma.cuh)."""
    r, k = w.shape
    return w.to(torch.bfloat16).view(r // 8, 8, k // 8, 8).permute(0, 2, 1, 3).contiguous()


@pytest.mark.parametrize("n,k", [(128, 128), (64, 64), (256, 128), (16, 32), (128, 272), (32, 144)])
def test_umma_gemm_stage(built_lib, n, k):
    from mazero_b200 import _lib

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(n * 1000 + k)
    a = torch.randn(128, k, generator=g).to(dev)
    w = (torch.randn(n, k, generator=g) / k ** 0.5).to(dev)
    out = torch.full((128, n), float("nan"), device=dev)
    wp = pack_operand(w)
    _lib.check(_lib.lib.maz_dbg_umma_gemm(a.data_ptr(), wp.data_ptr(), out.data_ptr(), n, k,
                                          C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    ref = a.to(torch.bfloat16).float() @ w.to(torch.bfloat16).float().t()   # bf16 operands, fp32 accumulate
    err = (out - ref).abs().max().item()
    assert err < 2e-3, f"max abs err {err}"     # fp32 accumulation-order noise only
