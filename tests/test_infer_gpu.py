"""GPU: the fused tensor-core inference path against a plain PyTorch fp32 reference."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


from mazero_b200.fused import pack_operand  # noqa: E402  (host-side packing under test as well)


@pytest.mark.parametrize("n,k", [(128, 128), (64, 64), (256, 128), (16, 32), (128, 272), (32, 144)])
def test_umma_gemm_stage(built_lib, n, k):
    from mazero_b200 import _lib

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(n * 1000 + k)
    a = torch.randn(128, k, generator=g).to(dev)
    w = (torch.randn(n, k, generator=g) / k ** 0.5).to(dev)
    out = torch.full((128, n), float("nan"), device=dev)
    wp = pack_operand(w).contiguous()
    _lib.check(_lib.lib.maz_dbg_umma_gemm(a.data_ptr(), wp.data_ptr(), out.data_ptr(), n, k,
                                          C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    ref = a.to(torch.bfloat16).float() @ w.to(torch.bfloat16).float().t()   # bf16 operands, fp32 accumulate
    err = (out - ref).abs().max().item()
    assert err < 2e-3, f"max abs err {err}"     # fp32 accumulation-order noise only


@pytest.mark.parametrize("kernel", ["twin", "tcgen05", "small"])
@pytest.mark.parametrize("N,A,B", [(3, 9, 100), (5, 11, 70), (10, 18, 33), (27, 36, 9), (1, 5, 130), (2, 3, 16), (2, 40, 20),
                                   (30, 17, 5), (3, 9, 1), (27, 36, 1300),    # 1300 x 27: 325 tiles = more than one wave of tile pairs
                                   (8, 12, 40), (16, 6, 10)])                # teams that fill a 32-row warp exactly
def test_fused_recurrent_inference_matches_fp32_reference(built_lib, N, A, B, kernel):
    """bf16 tensor-core kernels (tcgen05 128-row tiles; small-batch 32-row tiles on warp-level MMAs) vs the plain fp32 torch
    forward (same weights).  Tolerance: bf16 operand
    rounding (2^-8 relative) through ~20 chained GEMMs -> 2% of each tensor's dynamic range (max |ref|,
    at least 1); probabilities 1e-2 absolute."""
    from mazero_b200.inference import SmacInference
    from mazero_b200.synthetic import random_state_dict

    dev = torch.device("cuda:0")
    sd = random_state_dict(N, A, seed=N, head_scale=30.0)   # heads large enough to give non-uniform policies
    ref = SmacInference(sd, N, A, device=dev, mode="fp32")
    fus = SmacInference(sd, N, A, device=dev, mode="bf16")
    g = torch.Generator().manual_seed(B)
    S = 3
    pool = torch.randn(S, B, N * 128, generator=g).to(dev)
    idx = torch.randint(0, S, (B,), generator=g).to(torch.int32).to(dev)
    act = torch.randint(0, A, (B, N), generator=g).to(torch.int32).to(dev)
    h = pool.view(-1, N * 128).index_select(0, idx.long() * B + torch.arange(B, device=dev))
    nxt_ref, rew_ref, val_ref, log_ref = ref.recurrent(h, act)

    nxt = torch.full((B, N * 128), float("nan"), device=dev)
    rew = torch.full((B,), float("nan"), device=dev)
    val = torch.full((B,), float("nan"), device=dev)
    probs = torch.full((B, N, A), float("nan"), device=dev)
    beta = torch.full((B, N, A), float("nan"), device=dev)
    logits = torch.full((B, N, A), float("nan"), device=dev)
    greedy = torch.full((B, N), -1, dtype=torch.int32, device=dev)
    fus.recurrent_fused(B, pool, idx, act, nxt, rew, val, probs, beta, greedy, logits, kernel=kernel)
    torch.cuda.synchronize()
    for name, x, y, rel in (("next_hidden", nxt, nxt_ref, 2e-2), ("policy_logits", logits, log_ref, 2e-2),
                            ("reward", rew, rew_ref, 4e-2), ("value", val, val_ref, 4e-2),   # + the steep inv_h transform
                            ("probs", probs, torch.softmax(log_ref, -1), 1e-2)):
        assert torch.isfinite(x).all(), name
        err = (x - y).abs().max().item()
        tol = rel * max(1.0, y.abs().max().item())
        print(f"{name}: max abs err {err:.4g} (tol {tol:.3g}, ref range {y.abs().max().item():.3g})")
        assert err < tol, f"{name}: max abs err {err} > {tol}"
    assert torch.allclose(beta, probs / probs.sum(-1, keepdim=True), atol=1e-6)
    assert torch.allclose(probs.sum(-1), torch.ones(B, N, device=dev), atol=1e-5)
    # greedy = argmax of the kernel's own logits
    assert torch.equal(greedy.long(), logits.argmax(-1))
    # sequential mode: only agent `cur` is written, as (B,1,A)
    cur = N - 1
    p1 = torch.full((B, 1, A), float("nan"), device=dev)
    b1 = torch.full((B, 1, A), float("nan"), device=dev)
    fus.recurrent_fused(B, pool, idx, act, nxt, rew, val, p1, b1, None, None, tree_agents=1, cur=cur, inv_tau=0.5, kernel=kernel)
    torch.cuda.synchronize()
    assert torch.allclose(p1[:, 0], probs[:, cur], atol=1e-6)
    bt = probs[:, cur] ** 0.5
    assert torch.allclose(b1[:, 0], bt / bt.sum(-1, keepdim=True), atol=1e-4)
    # joint mode with a sampling temperature: every agent's beta = probs^(1/tau) renormalised (mcts_sampled.py:160)
    p2 = torch.full((B, N, A), float("nan"), device=dev)
    b2 = torch.full((B, N, A), float("nan"), device=dev)
    fus.recurrent_fused(B, pool, idx, act, nxt, rew, val, p2, b2, None, None, inv_tau=0.5, kernel=kernel)
    torch.cuda.synchronize()
    assert torch.allclose(p2, probs, atol=1e-6)
    bt = probs ** 0.5
    assert torch.allclose(b2, bt / bt.sum(-1, keepdim=True), atol=1e-4)


@pytest.mark.parametrize("N,A,B", [(27, 36, 300), (5, 11, 1000), (10, 18, 500)])
def test_twin_kernel_is_deterministic(built_lib, N, A, B):
    """Race guard for the two-tiles-in-flight kernel (compute-sanitizer is closed on this pool): its epilogue warps are not
    synchronised per stage any more (mbarrier hand-off, quadrant barriers, two CTA-wide barriers), so a shared-memory hazard
    would show as run-to-run differences.  200 launches on the same inputs, interleaved with launches of another size (different
    timing), must give bit-identical outputs."""
    from mazero_b200.inference import SmacInference
    from mazero_b200.synthetic import random_state_dict

    dev = torch.device("cuda:0")
    fus = SmacInference(random_state_dict(N, A, seed=7, head_scale=30.0), N, A, device=dev, mode="bf16")
    g = torch.Generator().manual_seed(B)
    pool = torch.randn(2, B, N * 128, generator=g).to(dev)
    idx = torch.zeros(B, dtype=torch.int32, device=dev)
    act = torch.randint(0, A, (B, N), generator=g).to(torch.int32).to(dev)

    def launch(b):
        out = [torch.zeros(b, N * 128, device=dev), torch.zeros(b, device=dev), torch.zeros(b, device=dev),
               torch.zeros(b, N, A, device=dev), torch.zeros(b, N, A, device=dev), torch.zeros(b, N, dtype=torch.int32, device=dev)]
        fus.recurrent_fused(b, pool[:, :b].contiguous() if b != B else pool, idx[:b], act[:b], *out[:5], out[5], None, kernel="twin")
        return out

    first = launch(B)
    torch.cuda.synchronize()
    for it in range(200):
        if it % 3 == 1:
            launch(max(1, B // 7))                    # a short launch in between: the next one starts on a differently warm GPU
        cur = launch(B)
        torch.cuda.synchronize()
        for a, b in zip(first, cur):
            assert torch.equal(a, b), f"launch {it} differs"
