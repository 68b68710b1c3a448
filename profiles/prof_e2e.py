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
