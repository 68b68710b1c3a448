"""Stage clocks (SM cycles, CTA 0 / thread 0) of the small-batch inference kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mazero_b200.inference import SmacInference  # noqa: E402
from mazero_b200.synthetic import WORKLOADS, random_state_dict  # noqa: E402

dev = torch.device("cuda:0")
name, B = os.environ.get("WL", "3m"), int(os.environ.get("ROOTS", "1024"))
N, A = WORKLOADS[name][:2]
inf = SmacInference(random_state_dict(N, A), N, A, device=dev, mode="bf16")
pool = torch.randn(2, B, N * 128, device=dev)
idx = torch.zeros(B, dtype=torch.int32, device=dev)
act = torch.randint(0, A, (B, N), device=dev, dtype=torch.int32)
r, v = torch.empty(B, device=dev), torch.empty(B, device=dev)
p, b = torch.empty(B, N, A, device=dev), torch.empty(B, N, A, device=dev)
clk = torch.zeros(64, dtype=torch.int64, device=dev)
for _ in range(3):
    inf.recurrent_fused(B, pool, idx, act, pool[1], r, v, p, b, kernel="small", dbg_clock=clk)
torch.cuda.synchronize()
c = clk.cpu().numpy()
names = ["gather", "params wait", "inproj"] + [f"L{l} {s}" for l in range(3) for s in ("qkv", "attention", "out-proj+LN", "linear1", "linear2+LN")] + \
        ["gather h", "dynamics (3 GEMM stages)", "heads1", "heads2"]
for i, n in enumerate(names):
    print(f"{n:28s} {c[i + 1] - c[i]:7d} cycles")
print("total", c[len(names)] - c[0], "cycles")
hs = ["gemms", "phase A + barrier", "B: graph sums", "policy: argmax", "policy: exp sums", "policy: writes", "barrier", "C: normalise + Y + barrier",
      "D: pool + barrier", "E: logits + barrier", "F: scalars"]
print("heads2 phases:", {n: int(c[41 + i] - c[40 + i]) for i, n in enumerate(hs)})
