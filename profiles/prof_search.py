"""Device-resident timing of one whole search through the native library (include/maz_search.h): persistent kernel vs the
library-built CUDA graph, per strategy / roots-per-CTA, CUDA events, L2 flushed between searches.
    python profiles/prof_search.py [workload] [reps] [strategies...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mazero_b200.inference import SmacInference  # noqa: E402
from mazero_b200.mcts_sampled import SampledMCTS, _DevicePlan, clear_caches  # noqa: E402
from mazero_b200.synthetic import WORKLOADS, NetworkOutput, SearchConfig, random_state_dict, root_hidden  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "3m"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    strategies = sys.argv[3:] or ["persistent", "graph"]
    N, A, B, S, K = WORKLOADS[wl]
    B = int(os.environ.get("MAZ_B", B))
    cur = os.environ.get("MAZ_CUR")
    cur = None if cur is None else int(cur)
    dev = torch.device("cuda:0")
    inf = SmacInference(random_state_dict(N, A, seed=0, head_scale=30.0), N, A, device=dev, mode="bf16")
    cfg = SearchConfig(A, S, K)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for strat in strategies:
        plan = _DevicePlan(inf, B, K, S, cur, cfg, 1.0, strategy=strat)
        times = []
        rpc = max(1, min(8, 32 // N, int(os.environ.get("MAZ_SEARCH_RPT", "8"))))
        ncta = (B + rpc - 1) // rpc
        clk = torch.zeros(2 * S + 4 * ncta + 64 + 4 * B + 32, dtype=torch.int64, device=dev)
        if plan.native is not None and plan.native.strategy == "persistent" and not os.environ.get("MAZ_NO_CLOCK"):
            plan.native.set_debug_clock(clk)      # (selects the kernel instance with the in-kernel cycle counters)
        for r in range(reps + 2):
            h = root_hidden(B, N, seed=r % 5).to(dev)
            pol, vlog = inf.prediction(h)
            val = inf._inv_transform(vlog, inf.vsup).reshape(B, 1)
            plan.stage_roots(h, torch.zeros(B, 1, device=dev), val, pol, None)
            rng = np.random.RandomState(r)
            noise = rng.dirichlet([0.3] * A, B * plan.Nt).astype(np.float32).reshape(B, plan.Nt, A)
            plan.inp_np["noise_raw0"][:] = noise
            plan._stream()
            plan._h2d(0, plan.in_turn[0][1])
            plan._apply_dev_fields()
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan._enqueue_search(0, cur, 11 + r, cfg, 0.25, 0)
            e1.record()
            plan.tree.check()
            if r >= 2:
                times.append(e0.elapsed_time(e1))
        t = np.array(times)
        call = clk.cpu().numpy()
        c = call[:2 * S].reshape(S, 2)
        per = call[2 * S:2 * S + 4 * ncta].reshape(-1, 4)
        per = per[per[:, 0] > 0]
        if plan.native is not None and plan.native.strategy == "persistent":
            assert plan.native.roots_per_cta() == rpc, (plan.native.roots_per_cta(), rpc)
        pt = call[2 * S + 4 * ncta + 64:][:4 * B].reshape(B, 4) if len(per) else np.zeros((0, 4))
        if len(pt) and pt[:, 0].any():
            order = np.argsort(-(pt[:, 0] + pt[:, 1]))
            print(f"  per tree (cycles/sim): expand+backup mean {pt[:, 0].mean() / S:.0f} max {pt[:, 0].max() / S:.0f}; select mean "
                  f"{pt[:, 1].mean() / S:.0f} max {pt[:, 1].max() / S:.0f}; mean depth {pt[:, 2].mean() / S:.2f}, max depth {pt[:, 3].max()}")
            print("  slowest trees [expand+backup/sim, select/sim, mean depth, max depth]: " +
                  "; ".join(f"{pt[i, 0] / S:.0f} {pt[i, 1] / S:.0f} {pt[i, 2] / S:.1f} {pt[i, 3]}" for i in order[:6]))
            d = pt[:, 2] / S
            for lo, hi in ((0, 2), (2, 3), (3, 5), (5, 8), (8, 100)):
                m = (d >= lo) & (d < hi)
                if m.any():
                    print(f"    mean depth in [{lo},{hi}): {m.sum()} trees, expand+backup {pt[m, 0].mean() / S:.0f}, select {pt[m, 1].mean() / S:.0f} cycles/sim")
        ts = call[2 * S + 4 * ncta + 64 + 4 * B:][:32]
        if ts[:6].all():
            nm = {0: "start", 1: "header", 2: "prefetch issued", 11: "beta/probs staged", 12: "cdf", 13: "sampled", 14: "keys merged",
                  15: "children", 3: "expansion done", 4: "backup", 5: "minmax+header"}
            order = [0, 1, 2, 11, 12, 13, 14, 15, 3, 4, 5]
            print("  tree 0, phase cycles per simulation (mean over the search): " +
                  ", ".join(f"{nm[b]} {ts[16 + i + 1] / S:.0f}" for i, b in enumerate(order[1:])))
        st_all = call[2 * S + 4 * ncta:][:64]
        hs = st_all[40:]
        hs = hs[hs > 0]
        if len(hs) > 2:
            print("  heads2 phases (cycles): " + " ".join(str(int(v)) for v in np.diff(hs)))
        st = st_all[:40]
        st = st[st > 0]
        if len(st) > 2:
            names = ["gather", "sync", "inproj"] + [f"L{l}.{x}" for l in range(3) for x in ("qkv", "attn", "out+ln", "lin1", "lin2+ln")] + \
                    ["-", "dynamics", "heads1", "heads2"]
            dl = np.diff(st)
            print("  stage cycles (CTA 0, middle simulation): " + ", ".join(f"{n} {int(v)}" for n, v in zip(names, dl)))
        if len(per):
            ghz = per[:, 0] / np.maximum(per[:, 1], 1)
            print(f"  {len(per)} CTAs: total cycles mean {per[:, 0].mean():.0f} min {per[:, 0].min()} max {per[:, 0].max()}; "
                  f"ns mean {per[:, 1].mean():.0f} max {per[:, 1].max()}; SM clock {ghz.mean():.3f} GHz; "
                  f"inference cycles/sim mean {per[:, 2].mean() / S:.0f} max {per[:, 2].max() / S:.0f}; "
                  f"tree cycles/sim mean {per[:, 3].mean() / S:.0f} min {per[:, 3].min() / S:.0f} max {per[:, 3].max() / S:.0f}")
        print(f"{wl} B={B} cur={cur} strategy={strat} ({plan.native.strategy if plan.native else 'legacy'}): "
              f"{t.mean():.3f} ms/search (min {t.min():.3f} max {t.max():.3f}) -> {B * S / t.mean() / 1e3:.2f} M root-sims/s"
              + (f"; CTA0 cycles/sim: inference {c[:, 0].mean():.0f} tree {c[:, 1].mean():.0f}" if c.any() else ""))
        out = {k: v.clone() for k, v in plan.turn_out[0].items()}
        plan.close()
        del plan
    clear_caches()


if __name__ == "__main__":
    main()
