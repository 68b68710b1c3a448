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
pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
from mazero_b200 import hostrng as _h
print("speculation:", _h.STATS)

# ---- wall-clock split of one search (host side) ------------------------------------------------------------------
import time
from mazero_b200 import hostrng
plan = next(iter(mcts._plans.values()))
acc = {}
def T(name, t0):
    acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0)
rs = np.random.RandomState(0)
for it in range(20):
    torch.cuda.synchronize()
    t = time.perf_counter(); plan.stage_roots(out.hidden_state, out.reward, out.value, out.policy_logits, None); T("stage_roots (hidden H2D enqueue + pinned staging)", t)
    t = time.perf_counter(); nz = hostrng.dirichlet_f32(rs, 0.3, A, B * N).reshape(B, N, A); sd = rs.choice(256); T("noise draw (parallel, bit-identical) + choice", t)
    t = time.perf_counter(); plan._stream(); plan.inp_np["noise_raw0"][:] = nz; plan._h2d(0, plan.in_turn[0][1]); T("staging H2D enqueue", t)
    t = time.perf_counter(); plan._enqueue_search(0, None, int(sd), cfg, 0.25, 0); T("enqueue root-prepare / reset / prepare / graph / readout", t)
    t = time.perf_counter(); plan.out_host.copy_(plan.out_flat, non_blocking=True); T("D2H enqueue", t)
    t = time.perf_counter(); plan.tree.check(); T("check (sync = GPU time)", t)
    t = time.perf_counter(); res = {k: v.copy() for k, v in plan.out_np.items()}; T("copy out arrays", t)
t = time.perf_counter()
for it in range(20):
    nz = rs.dirichlet([0.3] * A, B * N).astype(np.float32)
acc["(numpy dirichlet, for comparison)"] = time.perf_counter() - t
print({k: round(v / 20 * 1e3, 3) for k, v in acc.items()}, "ms per search")
