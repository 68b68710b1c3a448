"""Profiling driver: a few launches of the fused recurrent_inference kernel on the 3m workload shape.
    python profiles/prof_infer.py [workload] [launches]      (run plain first, then under ncu -k regex:k_recurrent)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mazero_b200 import build  # noqa: E402

build.build()
from mazero_b200.inference import SmacInference  # noqa: E402
from mazero_b200.synthetic import WORKLOADS, random_state_dict  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "3m"
n_launch = int(sys.argv[2]) if len(sys.argv) > 2 else 6
N, A, B, S, K = WORKLOADS[wl]
B = int(os.environ.get("ROOTS", B))
dev = torch.device("cuda:0")
inf = SmacInference(random_state_dict(N, A, seed=0), N, A, device=dev, mode="bf16")
g = torch.Generator().manual_seed(0)
pool = torch.randn(2, B, N * 128, generator=g).to(dev)
idx = torch.zeros(B, dtype=torch.int32, device=dev)
act = torch.randint(0, A, (B, N), generator=g).to(torch.int32).to(dev)
rew, val = torch.empty(B, device=dev), torch.empty(B, device=dev)
probs, beta = torch.empty(B, N, A, device=dev), torch.empty(B, N, A, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n_launch + 1)]
ev[0].record()
for i in range(n_launch):
    inf.recurrent_fused(B, pool, idx, act, pool[1], rew, val, probs, beta)
    ev[i + 1].record()
torch.cuda.synchronize()
print("k_recurrent_inference us per launch:", [round(ev[i].elapsed_time(ev[i + 1]) * 1e3, 1) for i in range(n_launch)])
print("finite:", bool(torch.isfinite(pool[1]).all()), bool(torch.isfinite(val).all()))

if os.environ.get("STAGE_CLOCKS"):
    dbg = torch.zeros(256, dtype=torch.int64, device=dev)
    inf.recurrent_fused(B, pool, idx, act, pool[1], rew, val, probs, beta, dbg_clock=dbg)
    torch.cuda.synchronize()
    t = dbg.cpu().numpy()
    sub = t[128:]
    t = t[:128]
    t = t[t > 0]
    print("stage clocks (cycles since first):", (t - t[0]).tolist())
    if sub.any():      # sub-phase clocks of one epilogue (two tiles x 16 entries), when the kernel was built with them
        for k in range(0, len(sub), 16):
            u = sub[k:k + 16]
            u = u[u > 0]
            if len(u):
                print("sub-phase cycles:", (u[1:] - u[:-1]).tolist())
