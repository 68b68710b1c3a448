"""CPU restatement (plain PyTorch, fp32) of the network the search calls -- TEST INFRASTRUCTURE.

Restates the SMAC `MAMuZeroNet` of the reference for the two methods on the search path:
  prediction(hidden)                     config/smac/model.py:494-501, 285-373
  recurrent_inference(hidden, action)    config/smac/model.py:562-574, 503-527, 198-282
  inverse value/reward transform         core/config.py:430-442, 463-499
with the SAME parameter names as the reference module tree, so a reference `state_dict()` loads with
`strict=True` (the golden fixtures tests/golden/model_*.npz carry such a state dict together with the
reference's outputs; tests/test_model_oracle.py pins this file against them).

It is the torch-fp32 reference the CUDA inference kernels are checked against, and the model used by the
CPU baseline leg of bench.py (the reference's Python cannot travel to the GPU box).  The product never
imports it.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def _mlp(sizes):
    """Linear -> LayerNorm -> ReLU ... -> Linear   (reference mlp(): config/smac/model.py:19-69)"""
    mods = []
    for i in range(len(sizes) - 1):
        mods.append(nn.Linear(sizes[i], sizes[i + 1]))
        if i < len(sizes) - 2:
            mods += [nn.LayerNorm(sizes[i + 1]), nn.ReLU()]
    return nn.Sequential(*mods)


class _GraphConv(nn.Module):  # config/smac/model.py:72-85
    def __init__(self, i, o):
        super().__init__()
        self.lin_layer = nn.Linear(i, o)


class _GraphNet(nn.Module):
    """2-layer all-ones-adjacency graph net, mean pool, linear head (config/smac/model.py:88-174)."""

    def __init__(self, sa_dim, n_agents, hidden, out_dim):
        super().__init__()
        self.sa_dim, self.n_agents = sa_dim, n_agents
        self.register_buffer("adj", torch.ones(n_agents, n_agents))
        self.gc1, self.nn_gc1 = _GraphConv(sa_dim, hidden), nn.Linear(sa_dim, hidden)
        self.gc2, self.nn_gc2 = _GraphConv(hidden, hidden), nn.Linear(hidden, hidden)
        self.V = nn.Linear(hidden, out_dim)

    def forward(self, x):
        b = x.shape[0]
        x = x.view(b, self.n_agents, self.sa_dim)
        adj = self.adj.unsqueeze(0).expand(b, -1, -1)
        f = F.relu(torch.bmm(adj, self.gc1.lin_layer(x)) + self.nn_gc1(x))
        f = F.layer_norm(f, [f.size(-1)])
        o = F.relu(torch.bmm(adj, self.gc2.lin_layer(f)) + self.nn_gc2(f))
        o = F.layer_norm(o, [o.size(-1)])
        return self.V(o.mean(dim=1))


class _PosEnc(nn.Module):  # config/smac/attention.py:6-28
    def __init__(self, d, n_pos):
        super().__init__()
        tab = np.array([[p / np.power(10000, 2 * (j // 2) / d) for j in range(d)] for p in range(n_pos)])
        tab[:, 0::2] = np.sin(tab[:, 0::2])
        tab[:, 1::2] = np.cos(tab[:, 1::2])
        self.register_buffer("pos_table", torch.FloatTensor(tab).unsqueeze(0))

    def forward(self, x):
        return x + self.pos_table[:, : x.size(1)]


class _AttnEnc(nn.Module):  # config/smac/attention.py:31-43
    def __init__(self, layers, d, hidden):
        super().__init__()
        self.pos_embed = _PosEnc(hidden, 30)
        self.encoder = nn.TransformerEncoder(
            nn.TransformerEncoderLayer(d_model=d, nhead=8, dim_feedforward=hidden, dropout=0.1), layers,
            enable_nested_tensor=False)

    def forward(self, x):
        return self.encoder(self.pos_embed(x).permute(1, 0, 2)).permute(1, 0, 2)


class _Dynamics(nn.Module):  # config/smac/model.py:198-282
    def __init__(self, n, h, a, dyn_layers, rsup, gnn_hidden):
        super().__init__()
        self.attention_stack = nn.Sequential(nn.Linear(h + a, h), nn.ReLU(), _AttnEnc(3, h, h))
        self.fc_dynamic = _mlp([2 * h + a] + list(dyn_layers) + [h])
        self.reward_predictor = _GraphNet(h + a, n, gnn_hidden, rsup)


class _Prediction(nn.Module):  # config/smac/model.py:285-373
    def __init__(self, n, h, a, pol_layers, vsup, gnn_hidden):
        super().__init__()
        self.value_predictor = _GraphNet(h, n, gnn_hidden, vsup)
        self.fc_policy = _mlp([h] + list(pol_layers) + [a])


class _Representation(nn.Module):  # config/smac/model.py:176-195 (use_feature_norm=True)
    def __init__(self, obs, h, layers):
        super().__init__()
        self.feature_norm = nn.LayerNorm(obs)
        self.mlp = _mlp([obs] + list(layers) + [h])

    def forward(self, x):
        return self.mlp(self.feature_norm(x))


def inverse_support_transform(logits, support_min, support_max):
    """softmax . support -> inv_h   (core/config.py:430-442, 463-475, 494-499)"""
    eps = 0.001
    p = torch.softmax(logits, dim=-1)
    sup = torch.arange(support_min, support_max + 1, dtype=torch.float32, device=logits.device)
    x = torch.sum(sup * p, dim=-1, keepdim=True)
    sign = torch.where(x < 0, -torch.ones_like(x), torch.ones_like(x))
    out = ((torch.sqrt(1 + 4 * eps * (torch.abs(x) + 1 + eps)) - 1) / (2 * eps)) ** 2 - 1
    out = sign * out
    out = torch.where(torch.isnan(out), torch.zeros_like(out), out)
    out = torch.where(torch.abs(out) < eps, torch.zeros_like(out), out)
    return out


class OracleMAMuZeroNet(nn.Module):
    """Search-path subset of the reference MAMuZeroNet (config/smac/model.py:397-574): dynamics + prediction.
    The representation / projection networks are not on the path; their parameters are accepted and ignored
    by `load_reference_state_dict`."""

    def __init__(self, num_agents, action_space_size, hidden_state_size=128, fc_dynamic_layers=(128, 128),
                 fc_policy_layers=(32,), reward_support=(-5, 5), value_support=(-5, 5), gnn_hidden=64, obs_size=None,
                 fc_representation_layers=None):
        super().__init__()
        if obs_size is not None:   # representation network (config/smac/model.py:176-195), only for initial_inference
            rl = list(fc_representation_layers if fc_representation_layers is not None else (hidden_state_size, hidden_state_size))
            self.representation_network = _Representation(obs_size, hidden_state_size, rl)
        self.num_agents, self.action_space_size, self.hidden = num_agents, action_space_size, hidden_state_size
        self.reward_support, self.value_support = reward_support, value_support
        rs = reward_support[1] - reward_support[0] + 1
        vs = value_support[1] - value_support[0] + 1
        self.dynamics_network = _Dynamics(num_agents, hidden_state_size, action_space_size, fc_dynamic_layers, rs, gnn_hidden)
        self.prediction_network = _Prediction(num_agents, hidden_state_size, action_space_size, fc_policy_layers, vs, gnn_hidden)

    def load_reference_state_dict(self, sd):
        own = self.state_dict()
        sub = {k: torch.as_tensor(np.asarray(v)) for k, v in sd.items() if k in own}
        missing = set(own) - set(sub)
        assert not missing, f"missing parameters: {sorted(missing)[:5]}"
        self.load_state_dict(sub, strict=True)
        return self

    def init_like_reference(self, seed=0):
        """Random init in the spirit of the reference constructor (orthogonal relu-gain MLPs, ~0 heads):
        synthetic weights for benchmarks -- there is no checkpoint to load."""
        g = torch.Generator().manual_seed(seed)
        for name, m in self.named_modules():
            if isinstance(m, nn.Linear):
                w = torch.empty_like(m.weight)
                nn.init.orthogonal_(w, gain=math.sqrt(2.0), generator=g)
                m.weight.data.copy_(w)
                m.bias.data.zero_()
        for head, scale in ((self.dynamics_network.reward_predictor.V, 3e-3),
                            (self.prediction_network.value_predictor.V, 3e-3),
                            (self.prediction_network.fc_policy[-1], 1e-3)):
            head.weight.data.copy_((torch.rand(head.weight.shape, generator=g) * 2 - 1) * scale)
            head.bias.data.zero_()
        return self

    @torch.no_grad()
    def initial_inference(self, observation):
        """config/smac/model.py:470-489, 542-559: (hidden (B,N*H), value (B,1), policy_logits (B,N,A))"""
        b = observation.shape[0]
        hidden = self.representation_network(observation.reshape(b * self.num_agents, -1)).reshape(b, -1)
        pol, vlog = self.prediction(hidden)
        return hidden, inverse_support_transform(vlog, *self.value_support), pol

    # -- the two methods the search calls ---------------------------------------------------------------
    def prediction(self, hidden):
        b = hidden.shape[0]
        value_logits = self.prediction_network.value_predictor(hidden)
        pol = self.prediction_network.fc_policy(hidden.view(b * self.num_agents, self.hidden))
        return pol.view(b, self.num_agents, self.action_space_size), value_logits

    def dynamics(self, hidden, action):
        b, n, h = hidden.shape[0], self.num_agents, self.hidden
        hs = hidden.view(b, n, h)
        onehot = F.one_hot(action.long(), num_classes=self.action_space_size).float()
        attn = self.dynamics_network.attention_stack(torch.cat([hs, onehot], dim=2))
        upd = self.dynamics_network.fc_dynamic(torch.cat([hs, onehot, attn], dim=2).reshape(b * n, -1))
        nxt = upd.view(b, n, h) + hs
        reward_logits = self.dynamics_network.reward_predictor(torch.cat([nxt, onehot], dim=2).reshape(b, -1))
        return nxt.reshape(b, -1), reward_logits

    @torch.no_grad()
    def recurrent_inference(self, hidden, action):
        """Returns (next_hidden (B,N*H), reward (B,1), value (B,1), policy_logits (B,N,A)) as torch tensors."""
        nxt, reward_logits = self.dynamics(hidden, action)
        policy_logits, value_logits = self.prediction(nxt)
        reward = inverse_support_transform(reward_logits, *self.reward_support)
        value = inverse_support_transform(value_logits, *self.value_support)
        return nxt, reward, value, policy_logits


# ---------------------------------------------------------------------------------------------------------
# The matrix-game network (BASELINE configs[0]): plain MLPs over the concatenated state of all agents.
def _matrix_mlp(sizes, value_out=False):
    """Linear -> ReLU -> LayerNorm per layer; `value_out` drops the last ReLU + LayerNorm
    (reference mlp(): config/matrix/model.py:15-51)."""
    mods = []
    for i in range(len(sizes) - 1):
        mods += [nn.Linear(sizes[i], sizes[i + 1]), nn.ReLU(), nn.LayerNorm(sizes[i + 1])]
    return nn.Sequential(*(mods[:-2] if value_out else mods))


class _MatrixDynamics(nn.Module):  # config/matrix/model.py:86-126
    def __init__(self, d, na, dyn_layers, rew_layers, rsup):
        super().__init__()
        self.fc_dynamic = _matrix_mlp([d + na] + list(dyn_layers) + [d])
        self.fc_reward = _matrix_mlp([d + na] + list(rew_layers) + [rsup], value_out=True)


class _MatrixPrediction(nn.Module):  # config/matrix/model.py:129-166
    def __init__(self, n, h, a, val_layers, pol_layers, vsup):
        super().__init__()
        self.fc_value = _matrix_mlp([n * h] + list(val_layers) + [vsup], value_out=True)
        self.fc_policy = _matrix_mlp([h] + list(pol_layers) + [a], value_out=True)


class OracleMatrixMuZeroNet(nn.Module):
    """Search-path subset of the reference's matrix-game MAMuZeroNet (config/matrix/model.py:208-368) with the
    reference's parameter names.  `reward_support` / `value_support` = (min, max); (0, 0) is the scalar head
    of use_vectorization=False (core/config.py:378-381), whose inverse transform is the identity (:488,495)."""

    def __init__(self, num_agents=2, action_space_size=3, hidden_state_size=64, fc_dynamic_layers=(64, 64),
                 fc_reward_layers=(32,), fc_value_layers=(32,), fc_policy_layers=(32,), reward_support=(-3, 3),
                 value_support=(-10, 10), obs_size=None, fc_representation_layers=(64, 64)):
        super().__init__()
        if obs_size is not None:   # representation network (config/matrix/model.py:54-83, use_feature_norm=False)
            self.representation_network = nn.Module()
            self.representation_network.mlp = _matrix_mlp([obs_size] + list(fc_representation_layers) + [hidden_state_size])
        self.num_agents, self.action_space_size, self.hidden_state_size = num_agents, action_space_size, hidden_state_size
        self.hidden = hidden_state_size
        self.reward_support, self.value_support = tuple(reward_support), tuple(value_support)
        rs, vs = reward_support[1] - reward_support[0] + 1, value_support[1] - value_support[0] + 1
        d = num_agents * hidden_state_size
        self.dynamics_network = _MatrixDynamics(d, num_agents * action_space_size, fc_dynamic_layers, fc_reward_layers, rs)
        self.prediction_network = _MatrixPrediction(num_agents, hidden_state_size, action_space_size, fc_value_layers,
                                                    fc_policy_layers, vs)

    load_reference_state_dict = OracleMAMuZeroNet.load_reference_state_dict

    def init_like_reference(self, seed=0):
        g = torch.Generator().manual_seed(seed)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                w = torch.empty_like(m.weight)
                nn.init.orthogonal_(w, gain=math.sqrt(2.0), generator=g)
                m.weight.data.copy_(w)
                m.bias.data.zero_()
        return self

    @torch.no_grad()
    def initial_inference(self, observation):  # config/matrix/model.py:323-329, 345-356
        b = observation.shape[0]
        hidden = self.representation_network.mlp(observation.reshape(b * self.num_agents, -1)).reshape(b, -1)
        pol, v = self.prediction(hidden)
        return hidden, self._inv(v, self.value_support), pol

    def prediction(self, hidden):  # config/matrix/model.py:163-166, 320-322
        b = hidden.shape[0]
        value = self.prediction_network.fc_value(hidden)
        pol = self.prediction_network.fc_policy(hidden.reshape(-1, self.hidden_state_size))
        return pol.reshape(b, self.num_agents, self.action_space_size), value

    def dynamics(self, hidden, action):  # config/matrix/model.py:334-345, 118-126
        b = hidden.shape[0]
        onehot = F.one_hot(action.long(), num_classes=self.action_space_size).float().view(b, -1)
        state = self.dynamics_network.fc_dynamic(torch.cat([hidden, onehot], dim=1)) + hidden
        return state, self.dynamics_network.fc_reward(torch.cat([state, onehot], dim=1))

    def _inv(self, x, sup):
        return x if sup[1] == sup[0] else inverse_support_transform(x, *sup)

    @torch.no_grad()
    def recurrent_inference(self, hidden, action):  # config/matrix/model.py:358-368
        nxt, reward = self.dynamics(hidden, action)
        policy_logits, value = self.prediction(nxt)
        return nxt, self._inv(reward, self.reward_support), self._inv(value, self.value_support), policy_logits
