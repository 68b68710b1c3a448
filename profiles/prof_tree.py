"""Profiling driver: the tree kernels alone (device-pointer step API, injected random network outputs).
    python profiles/prof_tree.py [workload] [sims]      (run plain first, then under ncu -k regex:k_select|k_expand)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from mazero_b200 import build  # noqa: E402

build.build()
from mazero_b200 import cytree  # noqa: E402
from mazero_b200.synthetic import WORKLOADS  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "3m"
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 16
N, A, B, S, K = WORKLOADS[wl]
B = int(os.environ.get("ROOTS", B))
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
probs = torch.softmax(torch.randn(sims + 1, B, N, A, generator=g) * 0.2, -1).to(dev).contiguous()
vals = (torch.randn(sims + 1, B, generator=g) * 0.1).to(dev)
rews = (torch.randn(sims + 1, B, generator=g) * 0.1).to(dev)
noise = torch.softmax(torch.randn(B, N, A, generator=g), -1).to(dev).contiguous()
t = cytree.Tree_batch(B, N, A, K, S, 0.01, 3, 0.75, 0.8, device=0)
t.set_puct(19652.0, 1.25)
t.set_stream(torch.cuda.current_stream().cuda_stream)
ix = torch.empty(B, dtype=torch.int32, device=dev)
iy = torch.empty(B, dtype=torch.int32, device=dev)
act = torch.empty(B, N, dtype=torch.int32, device=dev)
t.prepare(rews[0], vals[0], probs[0], probs[0], K, 0.25, noise)
torch.cuda._sleep(int(5e7))
evs = []
for s in range(sims):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    t.batch_selection_device(19652.0, 1.25, 0.99, ix, iy, act)
    e[1].record()
    t.batch_expansion_and_backup(s + 1, 0.99, K, rews[s + 1], vals[s + 1], probs[s + 1], probs[s + 1])
    e[2].record()
    evs.append(e)
t.check()
print("k_select us:", [round(e[0].elapsed_time(e[1]) * 1e3, 1) for e in evs])
print("k_expand_backup us:", [round(e[1].elapsed_time(e[2]) * 1e3, 1) for e in evs])
tot, sl, s1, s2 = t.stats()
print("mean depth", s1 / (B * sims), "nodes/tree", tot.mean())

if os.environ.get("STAGE_CLOCKS"):
    from mazero_b200._lib import lib, check
    import ctypes as C
    dbg = torch.zeros(64 + 8 * B, dtype=torch.int64, device=dev)
    check(lib.maz_tree_set_debug_clock(t._h, C.c_void_p(dbg.data_ptr())))
    t.expansion_backup_selection_device(sims + 1 if sims + 1 <= S else S, 0.99, K, rews[sims], vals[sims], probs[sims], probs[sims],
                                        19652.0, 1.25, ix, iy, act)
    t.check()
    c = dbg.cpu().numpy()
    names = {0: "start", 1: "hdr loaded", 11: "beta/probs staged", 12: "cdf built", 13: "sampled", 14: "dedup", 15: "children written",
             2: "prefetch issued", 3: "expanded", 4: "backup done", 5: "minmax+hdr written", 6: "select done"}
    order = [0, 1, 2, 11, 12, 13, 14, 15, 3, 4, 5, 6]
    print("fused tree kernel, tree 0, cycles since start:", {names[k]: int(c[k] - c[0]) for k in order})
    per = c[64:].reshape(B, 8)
    exp_c, bak_c, sel_c, len_b, len_s, pre = per[:, 0], per[:, 1], per[:, 2], per[:, 3], per[:, 4], per[:, 5]
    print("2-warp kernel, all trees: expansion cycles mean/max", exp_c.mean(), exp_c.max(), " backup mean/max", bak_c.mean(), bak_c.max(),
          " selection mean/max", sel_c.mean(), sel_c.max(), " until barrier mean/max", pre.mean(), pre.max())
    for d in sorted(set(len_b.tolist())):
        m = len_b == d
        print(f"  backup depth {d}: {m.sum()} trees, backup cycles mean {bak_c[m].mean():.0f}")
    for d in sorted(set(len_s.tolist())):
        m = len_s == d
        print(f"  selection depth {d}: {m.sum()} trees, selection cycles mean {sel_c[m].mean():.0f}")
