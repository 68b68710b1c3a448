"""Host-side profile of the public entry point (SampledMCTS.batch_search with host buffers)."""
import cProfile
import os
import pstats
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mazero_b200.inference import SmacInference  # noqa: E402
from mazero_b200.mcts_sampled import SampledMCTS  # noqa: E402
from mazero_b200.synthetic import NetworkOutput, SearchConfig, WORKLOADS, random_state_dict, root_hidden  # noqa: E402

N, A, B, S, K = WORKLOADS["3m"]
dev = torch.device("cuda:0")
cfg = SearchConfig(A, S, K)
inf = SmacInference(random_state_dict(N, A), N, A, device=dev, mode="bf16")
h = root_hidden(B, N, pinned=True)
pol, vlog = inf.prediction(h.to(dev))
out = NetworkOutput(h, np.zeros((B, 1), np.float32), inf._inv_transform(vlog, inf.vsup).cpu().numpy().reshape(B, 1), pol.cpu().numpy())
mcts = SampledMCTS(cfg, np.random.RandomState(1))
for _ in range(3):
    mcts.batch_search(inf, out, None, None, N, None, dev, add_noise=True)
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    mcts.batch_search(inf, out, None, None, N, None, dev, add_noise=True)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)

# ---- fine-grained wall-clock split of _DevicePlan.run ---------------------------------------------------------
import time
plan = next(iter(mcts._plans.values()))
probs = np.random.rand(B, N, A).astype(np.float32); beta = probs.copy(); noises = probs.copy()
r0 = np.zeros(B, np.float32); v0 = np.zeros(B, np.float32)
acc = {}
def T(name, t0):
    torch.cuda.synchronize() if name.endswith("*") else None
    acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0)
for it in range(20):
    torch.cuda.synchronize()
    t = time.perf_counter(); stream = torch.cuda.current_stream(dev); plan.tree.set_stream(stream.cuda_stream); T("set_stream", t)
    cp = lambda dst, src: dst.copy_(src if torch.is_tensor(src) else torch.from_numpy(np.ascontiguousarray(src)), non_blocking=True)
    t = time.perf_counter(); cp(plan.pool[0], h.reshape(B, -1)); T("h2d hidden (pinned)", t)
    t = time.perf_counter(); cp(plan.root_r, r0); cp(plan.root_v, v0); cp(plan.root_p, probs); cp(plan.root_b, beta); cp(plan.root_n, noises); T("h2d 5 small numpy", t)
    t = time.perf_counter(); plan.tree.reset(it, 0.01, 0.75, 0.8, 0); plan.tree.prepare(plan.root_r, plan.root_v, plan.root_p, plan.root_b, K, 0.25, plan.root_n); T("reset+prepare launch", t)
    t = time.perf_counter(); plan.graph.replay(); T("graph.replay launch", t)
    t = time.perf_counter(); plan.tree.readout_device(0.99, plan.out); plan.out_host.copy_(plan.out_flat, non_blocking=True); T("readout launch + d2h enqueue", t)
    t = time.perf_counter(); plan.tree.check(); T("check (sync = GPU time)", t)
    t = time.perf_counter(); res = {k: v.copy() for k, v in plan.out_np.items()}; T("copy out arrays", t)
rs = np.random.RandomState(0)
t = time.perf_counter()
for it in range(20):
    nz = rs.dirichlet([0.3] * A, B * N).astype(np.float32)
acc["numpy dirichlet"] = time.perf_counter() - t
print({k: round(v / 20 * 1e3, 3) for k, v in acc.items()}, "ms per search")
