"""The reference's unit-test fixtures, restated: MockConfig / MockModel of unit_test_mcts.py:41-130."""
from collections import namedtuple

import numpy as np
import torch

NetworkOutput = namedtuple("NetworkOutput", ["hidden_state", "reward", "value", "policy_logits"])


class MockConfig:
    def __init__(self, num_agents=2, action_space_size=3, num_simulations=3, sampled_action_times=5):
        self.pb_c_base, self.pb_c_init, self.discount = 19652.0, 1.25, 0.99
        self.mcts_rho, self.mcts_lambda = 0.75, 0.8
        self.root_dirichlet_alpha, self.root_exploration_fraction = 0.3, 0.25
        self.num_simulations = num_simulations
        self.action_space_size = action_space_size
        self.num_agents = num_agents
        self.sampled_action_times = sampled_action_times
        self.tree_value_stat_delta_lb = 0.01


class MockModel:
    """value .5 at the root, then .6; reward .1; zero logits everywhere; hidden += 0.01."""

    def __init__(self, num_agents, action_space_size, hidden_per_agent=4, device="cpu"):
        self.N, self.A, self.D, self.device = num_agents, action_space_size, num_agents * hidden_per_agent, device

    def eval(self):
        return self

    def initial_inference(self, batch):
        h = torch.rand((batch, self.D), device=self.device) * 0.1
        return NetworkOutput(h, np.zeros((batch, 1), np.float32), np.full((batch, 1), 0.5, np.float32),
                             np.zeros((batch, self.N, self.A), np.float32))

    def prediction(self, hidden):
        b = hidden.shape[0]
        return np.zeros((b, self.N, self.A), np.float32), np.zeros((b, 1), np.float32)

    def recurrent_inference(self, hidden, action):
        b = hidden.shape[0]
        assert tuple(action.shape) == (b, self.N)
        return NetworkOutput(hidden + 0.01, np.full((b, 1), 0.1, np.float32), np.full((b, 1), 0.6, np.float32),
                             np.zeros((b, self.N, self.A), np.float32))


def sequential_search(search_fn, cfg, model, batch=1, seed=123, legal=None):
    """The per-agent loop of unit_test_mcts.py:135-172 / selfplay_worker.py:196-257 (greedy pick)."""
    out0 = model.initial_inference(batch)
    chosen = np.full((batch, cfg.num_agents), -1, dtype=np.int32)
    outs = []
    for k in range(cfg.num_agents):
        factor = chosen[:, :k].copy() if k > 0 else None
        o = search_fn(model, out0, k, factor, cfg.num_agents, legal)
        outs.append(o)
        chosen[:, k] = np.argmax(o.marginal_visit_count[:, 0, :], axis=-1)
    return outs, chosen
