"""CPU restatement of the reference's search driver -- TEST INFRASTRUCTURE / CPU baseline.

`reference_batch_search` follows core/mcts/tree_search/mcts_sampled.py:34-200 line by line (root prep,
simulation loop with per-simulation host<->model round trips, 13 readouts) over a CPU tree
(`OracleTreeBatch(kind="reference")` = the reference's own C++, or kind="port") and any model with the
reference's `prediction` / `recurrent_inference` interface (`OracleMAMuZeroNet`, mock models).

`current_agent_idx=None` runs the upstream joint mode (tree agent_num = true_num_agents).
Used by tests (known answers), by bench.py's cpu_baseline leg and by `bench.py --impl reference`.
"""
from collections import namedtuple

import numpy as np
import torch

from .pyoracle import OracleTreeBatch

SearchOutput = namedtuple("SearchOutput", [
    "value", "marginal_visit_count", "marginal_priors", "sampled_actions", "sampled_visit_count", "sampled_pred_probs",
    "sampled_beta", "sampled_beta_hat", "sampled_priors", "sampled_imp_ratio", "sampled_pred_values",
    "sampled_mcts_values", "sampled_rewards", "sampled_qvalues"])

NetworkOutput = namedtuple("NetworkOutput", ["hidden_state", "reward", "value", "policy_logits"])


def _np(x):
    return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)


def _softmax(x):
    e = np.exp(x - np.max(x, axis=-1, keepdims=True))
    return e / np.sum(e, axis=-1, keepdims=True)


def model_recurrent(model, hidden, action):
    """Adapter: reference models return a NetworkOutput-like object, OracleMAMuZeroNet returns a tuple."""
    out = model.recurrent_inference(hidden, action)
    if isinstance(out, tuple) and not hasattr(out, "hidden_state"):
        nxt, r, v, pl = out
        return NetworkOutput(nxt, _np(r), _np(v), _np(pl))
    return NetworkOutput(out.hidden_state, _np(out.reward), _np(out.value), _np(out.policy_logits))


def reference_batch_search(config, np_random, model, network_output, current_agent_idx, factor, true_num_agents,
                           legal_actions_lst=None, device=None, add_noise=False, sampled_tau=1.0, tree_kind="reference",
                           trace=None):
    joint = current_agent_idx is None
    cur = current_agent_idx
    A, K = config.action_space_size, config.sampled_action_times
    Nt = true_num_agents if joint else 1
    B = network_output.hidden_state.shape[0]
    rewards, values = _np(network_output.reward), _np(network_output.value)
    all_logits = _np(network_output.policy_logits)
    logits = all_logits.reshape(B, Nt, A) if joint else all_logits[:, cur, :].reshape(B, 1, A)
    probs = _softmax(logits)
    noises = np_random.dirichlet([config.root_dirichlet_alpha] * A, B * Nt if joint else B).astype(np.float32).reshape(B, Nt, A)
    eps = config.root_exploration_fraction if add_noise else 0.0
    legal = None
    if legal_actions_lst is not None:
        legal = legal_actions_lst.reshape(B, Nt, A) if joint else legal_actions_lst[:, cur, :].reshape(B, 1, A)
        probs *= legal
        probs += legal * 1e-4
        probs = probs / np.sum(probs, axis=-1, keepdims=True)
        noises *= legal
        noises += legal * 1e-4
        noises = noises / np.sum(noises, axis=-1, keepdims=True)
    pool = [network_output.hidden_state]
    trees = OracleTreeBatch(B, Nt, A, K, config.num_simulations, config.tree_value_stat_delta_lb, np_random.choice(256),
                            config.mcts_rho, config.mcts_lambda, kind=tree_kind)
    beta = probs * (1 - eps) + noises * eps
    beta = beta ** (1 / sampled_tau)
    if legal is not None:
        beta *= legal
    beta = beta / np.sum(beta, axis=-1, keepdims=True)
    trees.prepare(rewards.reshape(B).astype(np.float32), values.reshape(B).astype(np.float32), probs.astype(np.float32),
                  beta.astype(np.float32), K, eps, noises.astype(np.float32))
    with torch.no_grad():
        if hasattr(model, "eval"):
            model.eval()
        for sim in range(config.num_simulations):
            ix, iy, acts = trees.batch_selection(config.pb_c_base, config.pb_c_init, config.discount)
            hs = torch.vstack([pool[x][y] for x, y in zip(ix, iy)])
            if joint:
                joint_action = acts.astype(np.int32)
            else:
                joint_action = np.zeros((B, true_num_agents), dtype=np.int32)
                if factor is not None:
                    for k in range(cur):
                        joint_action[:, k] = factor[:, k]
                joint_action[:, cur] = acts.squeeze()
                prior_logits = _np(model.prediction(hs)[0])
                for k in range(cur + 1, true_num_agents):
                    joint_action[:, k] = np.argmax(prior_logits[:, k, :], axis=-1)
            out = model_recurrent(model, hs, torch.from_numpy(joint_action).to(device))
            pl = out.policy_logits.reshape(B, Nt, A) if joint else out.policy_logits[:, cur, :].reshape(B, 1, A)
            p = _softmax(pl)
            b = p ** (1 / sampled_tau)
            b = b / np.sum(b, axis=-1, keepdims=True)
            pool.append(out.hidden_state)
            r32, v32 = out.reward.reshape(B).astype(np.float32), out.value.reshape(B).astype(np.float32)
            if trace is not None:
                trace.append((np.asarray(ix), acts.copy(), joint_action.copy()))
            trees.batch_expansion_and_backup(sim + 1, config.discount, K, r32, v32, p.astype(np.float32), b.astype(np.float32))
    d = config.discount
    return SearchOutput(trees.get_roots_values(), trees.get_roots_marginal_visit_count(), trees.get_roots_marginal_priors(),
                        trees.get_roots_sampled_actions(), trees.get_roots_sampled_visit_count(),
                        trees.get_roots_sampled_pred_probs(), trees.get_roots_sampled_beta(), trees.get_roots_sampled_beta_hat(),
                        trees.get_roots_sampled_priors(), trees.get_roots_sampled_imp_ratio(),
                        trees.get_roots_sampled_pred_values(), trees.get_roots_sampled_mcts_values(),
                        trees.get_roots_sampled_rewards(), trees.get_roots_sampled_qvalues(d))
