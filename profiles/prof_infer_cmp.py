"""Fused inference, per-launch time by batch: tcgen05 kernels (128-row tiles; "twin" = two tiles in flight per CTA, "tcgen05" = first
generation, one tile per CTA) vs the small-batch kernel (32-row tiles, warp-level MMAs)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mazero_b200.inference import SmacInference  # noqa: E402
from mazero_b200.synthetic import WORKLOADS, random_state_dict  # noqa: E402

dev = torch.device("cuda:0")
shapes = [("3m", b) for b in (256, 1024, 2048, 3072, 4096, 8192, 16384)] + [("2s3z", 4096), ("mmm2", 1024), ("mmm2", 8192), ("27m", 1024), ("27m", 16384)]
for name, B in shapes:
    N, A = WORKLOADS[name][:2]
    inf = SmacInference(random_state_dict(N, A), N, A, device=dev, mode="bf16")
    pool = torch.randn(2, B, N * 128, device=dev)
    idx = torch.zeros(B, dtype=torch.int32, device=dev)
    act = torch.randint(0, A, (B, N), device=dev, dtype=torch.int32)
    r, v = torch.empty(B, device=dev), torch.empty(B, device=dev)
    p, b = torch.empty(B, N, A, device=dev), torch.empty(B, N, A, device=dev)
    res = {}
    for kern in ("twin", "tcgen05", "small"):
        for _ in range(5):
            inf.recurrent_fused(B, pool, idx, act, pool[1], r, v, p, b, kernel=kern)
        torch.cuda.synchronize()
        torch.cuda._sleep(int(5e7))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            inf.recurrent_fused(B, pool, idx, act, pool[1], r, v, p, b, kernel=kern)
        e1.record()
        torch.cuda.synchronize()
        res[kern] = e0.elapsed_time(e1) / 50 * 1e3
    rpt = 32 // N
    print(f"{name:5s} B={B:6d} tiles tcgen05={-(-B // (4 * rpt)):5d} small={-(-B // rpt):5d}   twin {res['twin']:8.1f} us   tcgen05 (v1) {res['tcgen05']:8.1f} us   small {res['small']:8.1f} us", flush=True)
