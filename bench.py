#!/usr/bin/env python
"""bench.py -- MCTS simulations/second of the batched sampled-MCTS hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 3m] [--mode joint|seq] [--impl ours|reference]

A "step" is ONE whole search of the workload: B roots x S simulations (prepare -> S x [select -> gather
hidden -> recurrent_inference -> softmax/beta -> expand+backup] -> readout).  Default workload is
BASELINE.json configs[1]: SMAC 3m-shaped, 3 agents x 9 actions, 1024 roots x 50 sims, K = 10 sampled
joint actions, synthetic root hidden states and random-init weights of the reference architecture.

`value`   device-resident throughput: inputs already in HBM, outputs left in HBM; CUDA events.
`e2e`     the same metric through the public API a worker calls (`SampledMCTS.batch_search`) with HOST
          buffers: numpy root preparation, H2D of the root tensors from pinned memory, the search, D2H of
          all readouts.
`roofline`     the dominant kernel of the step (fused inference, ~75 % of GPU time) against the measured sustained
               bf16 peak; `roofline_tree`: the tree kernel (expand+backup+select) against the measured HBM copy peak.
`cpu_baseline` the reference CPU path (reference C++ tree compiled into oracle/_ref when available, the
          reference's Python loop restated, the same weights on torch-CPU) on the box's host cores.
`--impl reference` times that CPU path alone (the reference arm).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_JSON_FD = None     # set when fd 1 has been redirected (multi-rank runs)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


METRIC = "mcts_sims_per_sec"
UNIT = "root-sims/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="3m")
    ap.add_argument("--mode", default="joint", choices=["joint", "seq"])
    ap.add_argument("--roots", type=int, default=0, help="override roots per GPU")
    ap.add_argument("--sims", type=int, default=0, help="override simulations")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--inference", default="bf16", choices=["bf16", "fp32"], help="bf16 = fused tcgen05 kernel, fp32 = parity mode")
    return ap.parse_args()


def workload(args):
    from mazero_b200.synthetic import WORKLOADS

    N, A, B, S, K = WORKLOADS[args.workload]
    if args.roots:
        B = args.roots
    if args.sims:
        S = args.sims
    return N, A, B, S, K


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# algorithmic FLOPs of one recurrent_inference per root (SURVEY.md 8d / BASELINE.md section 4)
FLOP_PER_ROOT = {"3m": 2.61e6, "2s3z": 4.36e6, "mmm2": 8.81e6, "27m": 24.41e6}


def bf16_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["bf16_tflops_sustained"]) if os.path.exists(p) else 1400.0


class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                       "-lms", "20", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def samples(self):
        try:
            self.f.flush()
            return sum(1 for r in open(self.f.name) if r.strip())
        except Exception:
            return 0

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                out["sm_max_mhz"] = float(r[1])
                for n, v in zip(names, r[2:6]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
        out["reasons"] = sorted(reasons)
        return out


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_search(args, N, A, B, S, K, sd, hidden, steps, warmup):
    """The reference's CPU path: reference C++ tree (oracle/_ref) when built, else the C restatement; the
    reference's Python loop restated (oracle/search_oracle.py); the same weights on torch CPU."""
    import torch
    from mazero_b200.synthetic import SearchConfig
    from oracle import pyoracle
    from oracle.model_oracle import OracleMAMuZeroNet, inverse_support_transform
    from oracle.search_oracle import NetworkOutput, reference_batch_search

    pyoracle.build()
    kind = "reference" if pyoracle.available("reference") else "port"
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = OracleMAMuZeroNet(N, A).load_reference_state_dict(sd).eval()
    cfg = SearchConfig(A, S, K)
    hidden = hidden.float().cpu()
    with torch.no_grad():
        pol, vlog = model.prediction(hidden)
        value = inverse_support_transform(vlog, -5, 5)
    out0 = NetworkOutput(hidden, np.zeros((B, 1), np.float32), value.numpy(), pol.numpy())
    cur = None if args.mode == "joint" else 0
    rs = np.random.RandomState(1)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        reference_batch_search(cfg, rs, model, out0, cur, None, N, None, "cpu", add_noise=True, tree_kind=kind)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return float(np.mean(times)), kind, cores


def run_reference_arm(args):
    import torch
    from mazero_b200.synthetic import random_state_dict, root_hidden

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N, A, B, S, K = workload(args)
    # bounded sample of the workload: at most 1024 roots per step (the full 3m configuration)
    Bs = min(B, 1024)
    sd = random_state_dict(N, A, seed=0)
    hidden = root_hidden(Bs, N, seed=0)
    sec, kind, cores = cpu_reference_search(args, N, A, Bs, S, K, sd, hidden, args.steps, args.warmup)
    val = Bs * S / sec
    sample = f"{Bs} roots x {S} sims per step, {args.steps} steps, torch CPU {cores} threads, tree single-threaded"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}-shaped {N} agents x {A} actions, {Bs} roots x {S} sims, K={K}, {args.mode} mode",
                   "roots": Bs, "sims": S, "sampled_times": K, "mode": args.mode},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from mazero_b200 import build
    build.build()
    from mazero_b200.inference import SmacInference
    from mazero_b200.mcts_sampled import SampledMCTS
    from mazero_b200.synthetic import NetworkOutput, SearchConfig, random_state_dict, root_hidden

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- mazero_b200 has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the ONE JSON line: NCCL printf()s its version banner to fd 1 at communicator creation, so fd 1 is
        # pointed at stderr for the whole run and the JSON line goes to a saved copy of the original stdout
        sys.stdout.flush()
        global _JSON_FD
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)

    N, A, B, S, K = workload(args)
    cur = None if args.mode == "joint" else 0
    Nt = N if cur is None else 1
    cfg = SearchConfig(A, S, K)
    sd = random_state_dict(N, A, seed=0)
    inf = SmacInference(sd, N, A, device=dev, mode=args.inference)
    # Synthetic inputs: NSETS different batches of root hidden states, rotated per step with a rank offset.  The search time
    # depends on the inputs (the tree step ends with the deepest of the 1024 trees: 3.3 - 3.6 ms across batches), so one fixed
    # batch would make the single-GPU number a draw from that range and the max-over-ranks of a multi-GPU run a maximum of
    # such draws; with the rotation every rank does the same work on average, whatever the number of ranks.
    NSETS = int(os.environ.get("MAZ_BENCH_INPUT_SETS", "5"))
    base_seed = int(os.environ.get("MAZ_BENCH_SEED", "0"))
    root_off = rank * B                              # global root index of this rank's first root (tree RNG seeding)
    mcts = SampledMCTS(cfg, np.random.RandomState(1), use_cuda_graph=not args.no_graph)

    def api_step(net_out):
        return mcts.batch_search(inf, net_out, cur, None, N, None, dev, add_noise=True, root_index_offset=root_off)

    outs_host, outs_dev, snaps = [], [], []
    plan = None
    for k in range(NSETS):
        hidden_host = root_hidden(B, N, seed=base_seed + k, pinned=True)
        hidden_dev = hidden_host.to(dev)
        pol, vlog = inf.prediction(hidden_dev)
        value = inf._inv_transform(vlog, inf.vsup)
        oh = NetworkOutput(hidden_host, np.zeros((B, 1), np.float32), value.cpu().numpy().reshape(B, 1), pol.cpu().numpy())
        outs_host.append(oh)
        outs_dev.append(oh._replace(hidden_state=hidden_dev))
        first = api_step(outs_dev[-1])               # (k = 0: builds the plan + captures the CUDA graph)
        assert int(first.marginal_visit_count[0, 0].sum()) == S
        plan = next(iter(mcts._plans.values()))
        # device-resident copy of this batch's prepared roots: hidden state + the arrays Tree_batch.prepare takes
        snaps.append([t.clone() for t in (plan.pool[0], plan.root_r, plan.root_v, plan.root_p, plan.root_b, plan.root_n)])
    hidden_host = outs_host[0].hidden_state

    def load_set(k):
        for dst, src in zip((plan.pool[0], plan.root_r, plan.root_v, plan.root_p, plan.root_b, plan.root_n), snaps[k]):
            dst.copy_(src, non_blocking=True)

    # ---- device-resident step: reset + prepare + S simulations (graph) + readout kernel, nothing leaves HBM
    stream = torch.cuda.current_stream(dev)
    gather = None
    if world > 1 and not os.environ.get("MAZ_BENCH_NO_GATHER"):   # one exchange per search: root values + visit counts to every rank (learner = rank 0)
        gather = torch.empty(world * plan.out_flat.numel(), dtype=plan.out_flat.dtype, device=dev)
        comm = torch.cuda.Stream(dev)                      # the exchange of search i overlaps search i+1
        ev_ready, ev_done = torch.cuda.Event(), torch.cuda.Event()
        ev_done.record(stream)

    def dev_step(seed):
        load_set((seed + rank) % NSETS)                    # 1.6 MB of device-to-device copies, inside the timed step
        plan.tree.reset(seed, cfg.tree_value_stat_delta_lb, cfg.mcts_rho, cfg.mcts_lambda, root_off)
        plan.tree.prepare(plan.root_r, plan.root_v, plan.root_p, plan.root_b, K, cfg.root_exploration_fraction, plan.root_n)
        if plan.graph is not None:
            plan.graph.replay()
        else:
            plan._loop()
        if gather is not None:
            stream.wait_event(ev_done)                     # the previous search's readouts have been sent
        plan.tree.readout_device(cfg.discount, plan.out)
        if gather is not None:
            # ONE collective per search (NCCL over NVLink), on its own stream: it needs every rank to arrive, so run
            # synchronously it would add the rank skew of every 3 ms search to the step; here search i+1 hides it.
            # Each timed step contains the wait for the previous step's exchange, so n steps account for n exchanges.
            ev_ready.record(stream)
            with torch.cuda.stream(comm):
                comm.wait_event(ev_ready)
                dist.all_gather_into_tensor(gather, plan.out_flat)
                ev_done.record(comm)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def timed(fn, n, w):
        for i in range(w):
            fn(i)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        evs = []
        for i in range(n):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            fn(w + i)
            b.record(stream)
            evs.append((a, b))
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local) if rank == 0 else None
    ms_dev = timed(lambda i: dev_step(100 + i), args.steps, max(args.warmup, 3))
    # a default run times 10 x 3.6 ms: shorter than nvidia-smi's start-up.  Keep the GPU under the same load (untimed steps,
    # the same number on every rank: ms_dev is the all-reduced maximum) until ~0.4 s have passed, so that the clocks are
    # sampled under load and never missing.
    extra = int(max(0.0, 0.4 - args.steps * ms_dev * 1e-3) / (ms_dev * 1e-3)) + 1
    for i in range(extra):
        dev_step(1000 + i)
    torch.cuda.synchronize(dev)
    clocks = sampler.stop() if sampler else None
    plan.tree.check()

    # ---- end to end through the public API with host buffers (pinned hidden state, numpy everything else)
    def e2e_step(i):
        api_step(outs_host[(i + rank) % NSETS])

    # batch_search synchronises internally (it returns numpy); wall-clock == device time here, but keep events
    ms_e2e = timed(e2e_step, args.steps, max(args.warmup, 3))
    h2d = hidden_host.numel() * 4 + 4 * plan.in_turn[0][1]      # root hidden state + the pinned staging block
    d2h = plan.out_flat.numel() * 4

    # ---- sequential-agent mode: one environment step = N per-agent searches (selfplay_worker.py:196-257,
    # reanalyze_worker.py:278-327).  (a) the workers' loop over N batch_search calls with the host choosing each agent's
    # action in between; (b) SampledMCTS.search_agents: the same N searches back to back on the device, one sync.
    turns = None
    if cur is not None:
        def loop_step(i):
            acts = np.zeros((B, N), dtype=np.int32)
            for k in range(N):
                o = mcts.batch_search(inf, outs_host[(i + rank) % NSETS], k, acts[:, :k].copy() if k else None, N, None, dev, add_noise=True,
                                      root_index_offset=root_off)
                acts[:, k] = np.argmax(o.marginal_visit_count[:, 0, :], axis=-1)      # reanalyze_worker.py:309

        def fused_step(i):
            mcts.search_agents(inf, outs_host[(i + rank) % NSETS], N, None, dev, add_noise=True, turn="greedy", root_index_offset=root_off)

        fused_step(0)
        ms_loop = timed(loop_step, args.steps, max(args.warmup, 3))
        ms_fused = timed(fused_step, args.steps, max(args.warmup, 3))
        turns = {"unit": UNIT, "agents": N, "host_loop": world * B * S * N / (ms_loop * 1e-3), "host_loop_ms": ms_loop,
                 "search_agents": world * B * S * N / (ms_fused * 1e-3), "search_agents_ms": ms_fused,
                 "what": "one environment step = N per-agent searches; host_loop = N x batch_search with the action choice "
                         "on the host, search_agents = one device-resident call (one host sync)"}

    # ---- per-kernel timing of the tree kernels (eager loop, CUDA events around each launch) ---------------
    tot_nodes, last_len, sum_len, sum_exp = plan.tree.stats()
    dbar = sum_len / float(B * S)
    cbar = float((tot_nodes.sum() - B)) / float(sum_exp)
    plan.tree.reset(7, cfg.tree_value_stat_delta_lb, cfg.mcts_rho, cfg.mcts_lambda, root_off)
    plan.tree.prepare(plan.root_r, plan.root_v, plan.root_p, plan.root_b, K, cfg.root_exploration_fraction, plan.root_n)
    # the loop the CUDA graph runs: [k_recurrent_inference] -> [k_expand_backup_select]; timed eagerly behind a GPU sleep
    # so that the event pairs bracket back-to-back GPU execution and not host launch latency
    tree_ev, inf_ev = [], []
    torch.cuda._sleep(int(2.0e8))
    plan.tree.batch_selection_device(cfg.pb_c_base, cfg.pb_c_init, cfg.discount, plan.idx_x, plan.idx_y, plan.act)
    for s in range(S):
        joint = plan.act
        if cur is not None:
            flat = plan.idx_x.long() * B + plan.rows
            joint = torch.cat([plan.factor[:, :cur], plan.act, plan.greedy.view(-1, N).index_select(0, flat)[:, cur + 1:]],
                              dim=1).contiguous()
        if inf.fused is not None:
            ei = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ei[0].record(stream)
            inf.recurrent_fused(B, plan.pool, plan.idx_x, joint, plan.pool[s + 1], plan.sim_r, plan.sim_v, plan.sim_p, plan.sim_b,
                                plan.greedy[s + 1] if cur is not None else None, None, Nt, -1 if cur is None else cur, 1.0)
            ei[1].record(stream)
            inf_ev.append(ei)
            rew, val, p, bta = plan.sim_r, plan.sim_v, plan.sim_p, plan.sim_b
        else:
            flat = plan.idx_x.long() * B + plan.rows
            h = plan.pool.view(-1, N * inf.H).index_select(0, flat)
            _, rew, val, logits = inf.recurrent(h, joint, out_hidden=plan.pool[s + 1])
            if cur is not None:
                logits = logits[:, cur:cur + 1]
            p = torch.softmax(logits, dim=-1)
            bta = (p / p.sum(dim=-1, keepdim=True)).contiguous()
            p = p.contiguous()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e[0].record(stream)
        plan.tree.expansion_backup_selection_device(s + 1, cfg.discount, K, rew, val, p, bta, cfg.pb_c_base, cfg.pb_c_init,
                                                    plan.idx_x, plan.idx_y, plan.act)
        e[1].record(stream)
        tree_ev.append(e)
    torch.cuda.synchronize(dev)
    ms_tree = float(np.mean([a.elapsed_time(b) for a, b in tree_ev]))
    ms_inf = float(np.mean([a.elapsed_time(b) for a, b in inf_ev])) if inf_ev else None
    # algorithmic bytes per root-simulation of the tree step:
    #   expansion + backup: 8 + 8*N*A  reward, value, probs, beta in;  16 + C*(4N + 36)  leaf header + per-child fields;
    #                       40*(d+1)   visit, wsum/wtot r/w, value-log append, min-max entry r/w
    #   selection:          d*(16 + 20*C) parent header + per-child prior, visit, reward, wsum, wtot;  4*(2+N) idx, idy, act
    bytes_tree = B * (8 + 8 * Nt * A + 16 + cbar * (4 * Nt + 36) + 40 * (dbar + 1) + dbar * (16 + 20 * cbar) + 4 * (2 + Nt))
    hbm_peak, peak_src = peaks()
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and args.workload == "3m" and args.mode == "joint":
        traffic = json.load(open(tp))
    tree_roof = {"bound": "hbm", "kernel": "k_expand_backup_select2", "achieved": bytes_tree / (ms_tree * 1e-3) / 1e9,
                 "peak": hbm_peak, "unit": "GB/s", "frac": bytes_tree / (ms_tree * 1e-3) / 1e9 / hbm_peak,
                 "traffic": (traffic.get("k_expand_backup:3m:joint", 0) + traffic.get("k_select:3m:joint", 0)) or None,
                 "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_tree, "launch_ms": ms_tree}
    if ms_inf is not None:   # the dominant kernel of the step (~75 % of the GPU time): tensor-pipe roofline
        flop = B * FLOP_PER_ROOT.get(args.workload, 0.0)
        from mazero_b200 import fused as _fused
        small = _fused.use_small(B, N)
        roofline = {"bound": "tensor", "kernel": "k_recurrent_inference_small" if small else "k_recurrent_inference",
                    "achieved": flop / (ms_inf * 1e-3) / 1e12,
                    "peak": bf16_peak(), "unit": "TFLOP/s", "frac": flop / (ms_inf * 1e-3) / 1e12 / bf16_peak(),
                    "traffic": traffic.get("k_recurrent_inference_small:3m:joint" if small else "k_recurrent_inference:3m:joint"),
                    "peak_source": "measured sustained bf16 (MEASURED_PEAKS.json)",
                    "algorithmic_flop_per_launch": flop, "launch_ms": ms_inf,
                    "note": ("small-batch kernel: 32-row tiles on warp-level MMAs, one launch per simulation; bound by the dependent "
                             "stage chain of a tile (instruction issue / latency), not by the tensor pipe; see DESIGN.md section 5")
                            if small else "latency-bound 25-stage chain per 128-row tile; see DESIGN.md section 5"}
    else:
        roofline = tree_roof

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_sims = world * B * S
    line = {
        "metric": METRIC, "value": total_sims / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if inf.fused is not None else "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}-shaped {N} agents x {A} actions, {B} roots/GPU x {S} sims, K={K}, {args.mode} mode",
                   "roots_per_gpu": B, "sims": S, "sampled_times": K, "mode": args.mode, "tree_agents": Nt,
                   "inference": inf.mode, "cuda_graph": plan.graph is not None, "l2": "flushed (256 MiB memset) between timed steps",
                   "inputs": f"{NSETS} synthetic batches of root hidden states rotated per step (rank offset)",
                   "mean_search_depth": dbar, "mean_children": cbar, "parallelism": f"roots sharded x{world}",
                   **({"exchange": "one all_gather of the packed readouts per search, overlapped with the next search"} if world > 1 else {})},
        "e2e": {"value": total_sims / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": ms_e2e, "api": "SampledMCTS.batch_search (host numpy + pinned hidden state)"},
        # our kernels per search: k_seed, k_prepare, k_select (first simulation), per simulation the fused inference
        # kernel (bf16 mode) and the expand+backup(+next select) kernel, k_readout
        "gpu_launches": int(args.steps * (4 + (2 if inf.fused is not None else 1) * S)),
        "clocks": clocks,
        "roofline": roofline,
        "roofline_tree": tree_roof,
    }
    if turns is not None:
        line["e2e_agent_turns"] = turns
    if not args.no_cpu_baseline and world == 1:
        # bounded sample of the same workload on the host cores (one warm-up + two timed searches)
        Bs = min(B, 1024)
        sec, kind, cores = cpu_reference_search(args, N, A, Bs, S, K, sd, root_hidden(Bs, N, seed=0), 2, 1)
        line["cpu_baseline"] = {"value": Bs * S / sec, "unit": UNIT, "cores": cores, "kind": kind,
                                "sample": f"{Bs} roots x {S} sims, 2 timed searches after 1 warm-up; reference C++ tree "
                                          f"({kind}) single-threaded + torch CPU {cores} threads"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)
