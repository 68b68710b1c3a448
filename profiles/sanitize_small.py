"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck): tree step API + on-device search (fused
inference, both modes), small shapes.  `compute-sanitizer --tool memcheck python profiles/sanitize_small.py`"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _harness import MCTS, Inputs, drive  # noqa: E402
from mazero_b200 import cytree  # noqa: E402
from mazero_b200.inference import SmacInference  # noqa: E402
from mazero_b200.mcts_sampled import SampledMCTS  # noqa: E402
from mazero_b200.synthetic import NetworkOutput, SearchConfig, random_state_dict, root_hidden  # noqa: E402

for (B, N, A, K, S) in ((8, 3, 9, 10, 12), (4, 27, 36, 10, 6), (6, 1, 3, 5, 12)):
    inp = Inputs(B, N, A, S, seed=1, mode="random")
    out = drive(cytree.Tree_batch(B, N, A, K, S, MCTS["delta_lb"], 3, MCTS["rho"], MCTS["lam"]), inp, K)
    assert int(out["marginal_visit_count"][0, 0].sum()) == S
dev = torch.device("cuda:0")
for (B, N, A, K, S, cur) in ((8, 3, 9, 10, 8, None), (8, 3, 9, 5, 6, 1)):
    cfg = SearchConfig(A, S, K)
    inf = SmacInference(random_state_dict(N, A, seed=0), N, A, device=dev, mode="bf16")
    h = root_hidden(B, N).to(dev)
    pol, vlog = inf.prediction(h)
    out0 = NetworkOutput(h, np.zeros((B, 1), np.float32), inf._inv_transform(vlog, inf.vsup).cpu().numpy().reshape(B, 1), pol.cpu().numpy())
    factor = np.zeros((B, N), np.int32)
    res = SampledMCTS(cfg, np.random.RandomState(0), use_cuda_graph=False).batch_search(inf, out0, cur, factor, N, None, dev, add_noise=True)
    assert (res.marginal_visit_count.sum(axis=2) == S).all()
torch.cuda.synchronize()
print("sanitize_small ok")
