"""CPU restatement of the workers' per-agent turn logic -- TEST INFRASTRUCTURE.

  root_prepare           core/mcts/tree_search/mcts_sampled.py:57-106  (softmax, legal mask, noise mix, beta)
  select_action          core/utils.py:289-319
  eps_greedy_action      core/utils.py:322-334, with the two torch draws INJECTED (they depend on the mask only)
  selfplay_turns         core/selfplay_worker.py:196-257   (per-agent loop of DataWorker.run)
  reanalyze_turns        core/reanalyze_worker.py:278-345  (per-agent loop of _prepare_policy_re)

`select_action` / `eps_greedy_action` are pinned against the reference's own functions in
tests/test_reference_turns_cpu.py (build container) and by tests/golden/turns_kat.npz (generated from them).
The loops take `search_fn(agent_idx, factor) -> SearchOutput` so that they can run over any search
implementation; the product never imports this file.
"""
import numpy as np


def _softmax(x):
    e = np.exp(x - np.max(x, axis=-1, keepdims=True))
    return e / np.sum(e, axis=-1, keepdims=True)


def root_prepare(all_logits, cur, legal_actions_lst, noises, noise_eps, sampled_tau):
    """mcts_sampled.py:57-106.  all_logits (B,N,A); cur = agent index or None (joint); noises (B,Nt,A) raw Dirichlet
    draws (float32).  Returns probs, beta, noises (float32) exactly as Tree_batch.prepare receives them."""
    B, N, A = all_logits.shape
    Nt = N if cur is None else 1
    logits = all_logits.reshape(B, Nt, A) if cur is None else all_logits[:, cur, :].reshape(B, 1, A)
    probs = _softmax(logits)
    noises = noises.astype(np.float32).reshape(B, Nt, A).copy()
    legal = None
    if legal_actions_lst is not None:
        legal = legal_actions_lst.reshape(B, Nt, A) if cur is None else legal_actions_lst[:, cur, :].reshape(B, 1, A)
        probs *= legal
        probs += legal * 1e-4
        probs = probs / np.sum(probs, axis=-1, keepdims=True)
        noises *= legal
        noises += legal * 1e-4
        noises = noises / np.sum(noises, axis=-1, keepdims=True)
    beta = probs * (1 - noise_eps) + noises * noise_eps
    beta = beta ** (1 / sampled_tau)
    if legal is not None:
        beta *= legal
    beta = beta / np.sum(beta, axis=-1, keepdims=True)
    return probs.astype(np.float32), beta.astype(np.float32), noises.astype(np.float32)


def entropy_base2(pk):
    """scipy.stats.entropy(pk, base=2) (core/utils.py:318): renormalise, sum of entr(p) = -p log p, divide by log 2.
    scipy itself is called (it is in the image) so that the value is the reference's to the last bit."""
    from scipy.stats import entropy

    return entropy(pk, base=2)


def select_action(visit_counts, temperature=1, deterministic=True, uniform=None):
    """core/utils.py:289-319.  `uniform` is the ONE double np_random.choice(n, p=p) draws (random_sample())."""
    assert sum(visit_counts) > 0
    action_probs = [visit_count_i ** (1 / temperature) for visit_count_i in visit_counts]
    total_count = sum(action_probs)
    action_probs = np.array([x / total_count for x in action_probs])
    if deterministic:
        action_pos = np.argmax([v for v in visit_counts])
    else:
        cdf = action_probs.cumsum()             # RandomState.choice: cdf = p.cumsum(); cdf /= cdf[-1];
        cdf /= cdf[-1]                          #                     idx = cdf.searchsorted(uniform, side='right')
        action_pos = int(cdf.searchsorted(uniform, side="right"))
    return action_pos, entropy_base2(action_probs)


def eps_greedy_action(greedy_action, eps, eps_u, random_action):
    """core/utils.py:322-334 with injected draws: eps_u = torch.rand_like(...), random_action = Categorical(mask).sample()."""
    pick_random = int(eps_u < eps)
    return pick_random * int(random_action) + (1 - pick_random) * int(greedy_action)


def visit_policy(marginal_visits, legal_actions, A):
    """reanalyze_worker.py:319-327 (and the same quotient in selfplay_worker.py:283-286)."""
    if np.sum(marginal_visits) > 0:
        return marginal_visits / np.sum(marginal_visits)
    num_legal = np.sum(legal_actions)
    if num_legal > 0:
        return legal_actions / num_legal
    dummy = np.zeros(A)
    dummy[0] = 1.0
    return dummy


def selfplay_turns(search_fn, np_random, B, N, A, legal_actions_lst, temperature, greedy_epsilon, eps_u=None, random_action=None):
    """selfplay_worker.py:196-257.  search_fn(agent_idx, factor) must itself consume np_random like batch_search does.
    Returns (actions (B,N) int32, entropies (B,N), per-agent outputs)."""
    actions = np.full((B, N), -1, dtype=np.int32)
    entropies = np.zeros((B, N))
    outs = []
    for agent_idx in range(N):
        factor = actions[:, :agent_idx].copy() if agent_idx > 0 else None
        out = search_fn(agent_idx, factor)
        outs.append(out)
        for b in range(B):
            sampled_actions, counts = out.sampled_actions[b], out.sampled_visit_count[b]
            assert sampled_actions.size
            pos, ent = select_action(counts, temperature=temperature, deterministic=False, uniform=np_random.random_sample())
            a = sampled_actions[pos, 0]
            if eps_u is not None:
                a = eps_greedy_action(a, greedy_epsilon, eps_u[agent_idx][b], random_action[agent_idx][b])
            actions[b, agent_idx] = a
            entropies[b, agent_idx] = ent
    return actions, entropies, outs


def reanalyze_turns(search_fn, B, N, A, legal_actions_lst):
    """reanalyze_worker.py:278-345.  Returns (actions (B,N) int32, policy_dist (B,N,A) f64, probs (B,) f64, outputs)."""
    current_actions = np.zeros((B, N), dtype=np.int32)
    dist = np.zeros((B, N, A))
    outs = []
    for agent_idx in range(N):
        factor = current_actions[:, :agent_idx].copy() if agent_idx > 0 else None
        out = search_fn(agent_idx, factor)
        outs.append(out)
        for s in range(B):
            marginal = out.marginal_visit_count[s, 0, :]
            legal = legal_actions_lst[s, agent_idx, :] if legal_actions_lst is not None else np.ones(A)
            assert out.sampled_actions[s].size > 0 and np.sum(out.sampled_visit_count[s]) > 0 and np.sum(marginal) > 0
            current_actions[s, agent_idx] = np.argmax(marginal * legal)
            dist[s, agent_idx] = visit_policy(marginal, legal, A)
    probs = np.zeros(B)
    for s in range(B):
        prob_prod = 1.0
        for k in range(N):
            prob_prod *= dist[s, k][current_actions[s, k]]
        probs[s] = prob_prod
    return current_actions, dist, probs, outs
