"""GPU: the on-device search at BASELINE.json's FULL sizes, through size-independent properties (the CPU oracle cannot
replay 16 384 roots x 27 agents in test time):

* conservation: every simulation visits exactly one root child -> visit counts sum to S per root (and per agent in the
  marginals); the marginal visit counts are the scatter-add of the sampled visit counts over the sampled actions;
  sampled joint actions are distinct, in range, 1 <= #children <= K;
* independence (the property the multi-GPU sharding rests on): the first roots searched ALONE give bit-identical
  readouts to the same roots inside the full batch (same inputs, same seed, same global root index);
* determinism: the same search twice gives identical bits;
* 3m / 2s3z full sizes additionally replay EVERY simulation's recorded network outputs through the CPU oracle tree
  (bit-exact selections, visit counts, values, Q)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

FULL = {   # BASELINE.json configs[1..4]: (agents, actions, roots, sims, K)
    "3m": (3, 9, 1024, 50, 10),
    "2s3z": (5, 11, 4096, 100, 10),
    "mmm2": (10, 18, 8192, 50, 10),
    "27m": (27, 36, 16384, 50, 10),
}


def _setup(name, B=None):
    from mazero_b200.inference import SmacInference
    from mazero_b200.synthetic import SearchConfig, random_state_dict

    N, A, Bf, S, K = FULL[name]
    B = Bf if B is None else B
    inf = SmacInference(random_state_dict(N, A, seed=1, head_scale=30.0), N, A, device=DEV, mode="bf16")
    g = torch.Generator().manual_seed(7)
    hidden = torch.randn(Bf, N * 128, generator=g)[:B].to(DEV)
    pol, vlog = inf.prediction(hidden)
    value = inf._inv_transform(vlog, inf.vsup).cpu().numpy().reshape(B, 1)
    rng = np.random.RandomState(3)
    noises = rng.dirichlet([0.3] * A, Bf * N).astype(np.float32).reshape(Bf, N, A)[:B]
    legal = (rng.rand(Bf, N, A) < 0.8).astype(np.float32)[:B]
    legal[..., 1] = 1
    return inf, SearchConfig(A, S, K), hidden, np.zeros((B, 1), np.float32), value, pol.cpu().numpy(), legal, noises, (N, A, B, S, K)


def _run(inf, cfg, hidden, rewards, value, logits, legal, noises, dims, seed=11, record=None, graph=True):
    from mazero_b200.mcts_sampled import _DevicePlan

    N, A, B, S, K = dims
    plan = _DevicePlan(inf, B, K, S, None, cfg, 1.0, use_graph=graph)
    plan.record = record
    plan.stage_roots(hidden, rewards, value, logits, legal)
    out = plan.run(seed, cfg, cfg.root_exploration_fraction, noises, 0, cur=None)
    return out, plan


def _check_invariants(r, dims):
    N, A, B, S, K = dims
    n = r["num_children"]
    assert n.min() >= 1 and n.max() <= K
    vis, act = r["visit_count"], r["actions"]
    live = np.arange(K)[None, :] < n[:, None]
    assert (vis[~live] == 0).all() and (vis >= 0).all()
    assert (vis.sum(1) == S).all(), "every simulation passes through exactly one root child"
    assert (r["marginal_visit_count"].sum(2) == S).all()
    assert ((act >= 0) & (act < A)).all()
    # marginal visits = scatter-add of the sampled visits over each agent's sampled action
    marg = np.zeros((B, N, A), np.int64)
    bi = np.repeat(np.arange(B), K * N)
    ni = np.tile(np.arange(N), B * K)
    np.add.at(marg, (bi, ni, act.reshape(-1)), np.repeat(vis.reshape(-1), N))
    assert np.array_equal(marg, r["marginal_visit_count"])
    # sampled joint actions of a root are distinct (cnode.cpp:263-275: one child per distinct hashed joint action)
    key = (act.astype(np.int64) * (A ** np.arange(N))[None, None, :]).sum(2) if A ** N < 2 ** 62 else None
    if key is not None:
        key = np.where(live, key, -1 - np.arange(K)[None, :])
        srt = np.sort(key, axis=1)
        assert (srt[:, 1:] != srt[:, :-1]).all()
    for f in ("value", "qvalues", "mcts_values", "priors", "beta", "pred_probs"):
        assert np.isfinite(r[f]).all(), f
    assert (r["beta_hat"][live] > 0).all() and (r["beta_hat"][live] <= 1).all()
    np.testing.assert_allclose(r["beta_hat"].sum(1), 1.0, rtol=1e-5)      # counts / K over the distinct samples


@pytest.mark.parametrize("name", list(FULL))
def test_full_size_conservation_independence_determinism(built_lib, name):
    args = _setup(name)
    dims = args[-1]
    N, A, B, S, K = dims
    full, plan = _run(*args)
    _check_invariants(full, dims)
    again = plan.run(11, args[1], args[1].root_exploration_fraction, args[7], 0, cur=None)       # same plan, graph replay
    for k in full:
        assert np.array_equal(full[k], again[k]), f"determinism: {k}"
    # independence: the first roots searched alone (same kernel family as the full batch: forced)
    import os
    from mazero_b200 import fused

    Bs = max(32 // N, 1) * 13 + 1           # not a multiple of any tile size
    os.environ["MAZ_INFER_KERNEL"] = "small" if fused.use_small(B, N) else "tcgen05"
    os.environ["MAZ_INFER_TC"] = "twin" if fused.use_twin(B, N) else "v1"       # (and the same tcgen05 kernel generation)
    try:
        sub_args = _setup(name, B=Bs)
        sub, _ = _run(*sub_args)
    finally:
        del os.environ["MAZ_INFER_KERNEL"]
        del os.environ["MAZ_INFER_TC"]
    for k in full:
        assert np.array_equal(full[k][:Bs], sub[k]), f"independence: {k}"
    del plan
    torch.cuda.empty_cache()


@pytest.mark.parametrize("name", ["3m", "2s3z"])
def test_full_size_search_replays_bit_exact_through_oracle_tree(built_lib, oracle_built, name):
    args = _setup(name)
    inf, cfg, hidden, rewards, value, logits, legal, noises, dims = args
    N, A, B, S, K = dims
    rec = []
    out, plan = _run(*args, record=rec, graph=False)
    assert len(rec) == S
    orc = oracle_built.OracleTreeBatch(B, N, A, K, S, cfg.tree_value_stat_delta_lb, 11, cfg.mcts_rho, cfg.mcts_lambda)
    orc.prepare(plan.root_r.cpu().numpy(), plan.root_v.cpu().numpy(), plan.root_p.cpu().numpy(), plan.root_b.cpu().numpy(),
                K, cfg.root_exploration_fraction, plan.root_n.cpu().numpy())
    for s, (rew, val, p, b, ix, act) in enumerate(rec):
        oix, _, oact = orc.batch_selection(cfg.pb_c_base, cfg.pb_c_init, cfg.discount)
        assert np.array_equal(np.asarray(oix, np.int32), ix.cpu().numpy()), f"sim {s}: hidden_state_index_x"
        assert np.array_equal(oact, act.cpu().numpy()), f"sim {s}: selected actions"
        orc.batch_expansion_and_backup(s + 1, cfg.discount, K, rew.cpu().numpy(), val.cpu().numpy(), p.cpu().numpy(), b.cpu().numpy())
    ro = orc.readout(cfg.discount)
    assert np.array_equal(out["value"], orc.get_roots_values())
    assert np.array_equal(out["marginal_visit_count"], orc.get_roots_marginal_visit_count())
    for k in ("num_children", "actions", "visit_count", "qvalues", "mcts_values", "priors"):
        assert np.array_equal(out[k], ro[k]), k
    graphed, _ = _run(*args)                 # CUDA-graph replay == eager loop
    for k in out:
        assert np.array_equal(out[k], graphed[k]), k
