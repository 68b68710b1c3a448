"""Device-resident forward of the reference's MLP-family network for the search: `prediction` and
`recurrent_inference` of the matrix-game `MAMuZeroNet` (config/matrix/model.py:320-368; BASELINE configs[0]).

`MlpInference.from_model(model)` reads a reference-style module tree by parameter name
(dynamics_network.fc_dynamic.* / fc_reward.*, prediction_network.fc_value.* / fc_policy.*); the layer
chain -- Linear, ReLU, LayerNorm (config/matrix/model.py:44-49) -- is recovered from the state dict's
Sequential indices.  The fused path is one fp32 CUDA kernel per simulation (csrc/maz_mlp.cu, C ABI
`maz_mlp_recurrent`, include/maz_infer.h); `recurrent` / `prediction` are the same maths in plain torch ops
(used for the roots and as the floating-point cross-check in the tests).
"""
import ctypes as C

import torch
import torch.nn.functional as F

from ._lib import check, lib

MAXL, MAXW = 6, 640
LINEAR, RELU_LN, LN_RELU, LN_ONLY = 0, 1, 2, 3


class MlpLayer(C.Structure):
    _fields_ = [("wt", C.c_void_p), ("b", C.c_void_p), ("ln_w", C.c_void_p), ("ln_b", C.c_void_p),
                ("n_in", C.c_int), ("n_out", C.c_int), ("kind", C.c_int)]


class MlpNet(C.Structure):
    _fields_ = [("n", C.c_int), ("l", MlpLayer * MAXL)]


class MlpDesc(C.Structure):
    _fields_ = [("B", C.c_int), ("N", C.c_int), ("A", C.c_int), ("H", C.c_int), ("Nt", C.c_int), ("cur", C.c_int),
                ("inv_tau", C.c_float),
                ("pool", C.c_void_p), ("idx_x", C.c_void_p), ("actions", C.c_void_p), ("next_hidden", C.c_void_p),
                ("reward", C.c_void_p), ("value", C.c_void_p), ("probs", C.c_void_p), ("beta", C.c_void_p),
                ("greedy", C.c_void_p), ("logits_out", C.c_void_p),
                ("dyn", MlpNet), ("rew", MlpNet), ("val", MlpNet), ("pol", MlpNet),
                ("reward_support_min", C.c_int), ("reward_support_size", C.c_int),
                ("value_support_min", C.c_int), ("value_support_size", C.c_int),
                ("factor", C.c_void_p), ("greedy_pool", C.c_void_p)]


def _chain(sd, prefix):
    """[(linear_key_index, kind, ln_key_index or None)] of an nn.Sequential built by the reference's mlp()."""
    idx = sorted({int(k[len(prefix):].split(".")[0]) for k in sd if k.startswith(prefix)})
    lin = [i for i in idx if sd[f"{prefix}{i}.weight"].dim() == 2]
    out = []
    for i in lin:
        if (i + 2) in idx and (i + 2) not in lin and (i + 1) not in idx:
            out.append((i, RELU_LN, i + 2))       # Linear, ReLU, LayerNorm   (config/matrix/model.py:44-46)
        elif (i + 1) in idx and (i + 1) not in lin:
            out.append((i, LN_RELU, i + 1))       # Linear, LayerNorm, ReLU   (config/smac/model.py:52-58)
        else:
            out.append((i, LINEAR, None))
    return out


def supported(sd):
    keys = ("dynamics_network.fc_dynamic.0.weight", "dynamics_network.fc_reward.0.weight",
            "prediction_network.fc_value.0.weight", "prediction_network.fc_policy.0.weight")
    return all(k in sd for k in keys)


class _Net:
    def __init__(self, sd, prefix, device, registry):
        self.layers = []      # (w (out,in), wt (in,out), b, ln_w, ln_b, kind)
        self.keys = []
        for i, kind, j in _chain(sd, prefix):
            def g(k):
                t = sd[k].detach().to(device=device, dtype=torch.float32).contiguous()
                registry[k] = t
                return t
            w, b = g(f"{prefix}{i}.weight"), g(f"{prefix}{i}.bias")
            lw = lb = None
            if j is not None:
                lw, lb = g(f"{prefix}{j}.weight"), g(f"{prefix}{j}.bias")
            self.layers.append([w, w.t().contiguous(), b, lw, lb, kind])
        if not 1 <= len(self.layers) <= MAXL or any(L[0].shape[0] > MAXW for L in self.layers):
            raise RuntimeError(f"{prefix}: layer chain not supported by the MLP kernel")

    def retranspose(self):
        for L in self.layers:
            L[1].copy_(L[0].t())

    def __call__(self, x):
        for w, _, b, lw, lb, kind in self.layers:
            x = F.linear(x, w, b)
            if kind == RELU_LN:
                x = F.layer_norm(F.relu(x), [x.size(-1)], lw, lb)
            elif kind == LN_RELU:
                x = F.relu(F.layer_norm(x, [x.size(-1)], lw, lb))
        return x

    def fill(self, net: MlpNet):
        net.n = len(self.layers)
        for l, (w, wt, b, lw, lb, kind) in enumerate(self.layers):
            L = net.l[l]
            L.wt, L.b = wt.data_ptr(), b.data_ptr()
            L.ln_w = lw.data_ptr() if lw is not None else None
            L.ln_b = lb.data_ptr() if lb is not None else None
            L.n_out, L.n_in, L.kind = w.shape[0], w.shape[1], kind

    @property
    def out_features(self):
        return self.layers[-1][0].shape[0]


class RowMlp:
    """The representation network of `initial_inference` (config/smac/model.py:176-195: LayerNorm of the observation, then
    mlp(); config/matrix/model.py:54-83) as one launch of the row-wise MLP kernel (`maz_mlp_forward`), one row per agent
    observation; `torch_forward` is the same maths in torch ops."""

    def __init__(self, sd, device, registry, prefix="representation_network."):
        self.device = torch.device(device)
        self.net = _Net(sd, prefix + "mlp.", self.device, registry)
        self.pre = None
        if prefix + "feature_norm.weight" in sd:
            g = lambda k: sd[prefix + k].detach().to(device=self.device, dtype=torch.float32).contiguous()
            self.pre = (g("feature_norm.weight"), g("feature_norm.bias"))
            registry[prefix + "feature_norm.weight"], registry[prefix + "feature_norm.bias"] = self.pre
        self.in_features = self.net.layers[0][0].shape[1]
        self.out_features = self.net.out_features
        if len(self.net.layers) + (self.pre is not None) > MAXL or self.in_features > MAXW:
            raise RuntimeError("representation network not supported by the row-wise MLP kernel")

    def retranspose(self):
        self.net.retranspose()

    def torch_forward(self, x):
        if self.pre is not None:
            x = F.layer_norm(x, [x.size(-1)], self.pre[0], self.pre[1])
        return self.net(x)

    def __call__(self, x, stream=None):
        """x (rows, obs) fp32 CUDA tensor -> (rows, hidden)"""
        x = x.contiguous()
        rows = x.shape[0]
        y = torch.empty(rows, self.out_features, dtype=torch.float32, device=self.device)
        net = MlpNet()
        o = 0
        if self.pre is not None:
            L = net.l[0]
            L.wt = L.b = None
            L.ln_w, L.ln_b = self.pre[0].data_ptr(), self.pre[1].data_ptr()
            L.n_in = L.n_out = self.in_features
            L.kind = LN_ONLY
            o = 1
        tmp = MlpNet()
        self.net.fill(tmp)
        for l in range(tmp.n):
            for f, _ in MlpLayer._fields_:
                setattr(net.l[o + l], f, getattr(tmp.l[l], f))
        net.n = tmp.n + o
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        check(lib.maz_mlp_forward(C.byref(net), C.c_void_p(x.data_ptr()), int(rows), C.c_void_p(y.data_ptr()),
                                  C.c_void_p(s.cuda_stream)))
        return y


def initial_inference_device(inf, rep, observation):
    """`initial_inference` (config/smac/model.py:542-559, config/matrix/model.py:345-356) without leaving the device:
    representation network (row-wise MLP kernel), prediction heads, inverse value transform.  Returns
    (hidden (B, N*H), reward zeros (B, 1), value (B, 1), policy_logits (B, N, A)) as CUDA tensors, no synchronisation."""
    B = observation.shape[0]
    x = observation.reshape(B * inf.N, -1).to(device=inf.device, dtype=torch.float32)
    hidden = rep(x).view(B, inf.N * inf.H)
    pol, vlog = inf.prediction(hidden)
    value = inf._inv_transform(vlog, inf.vsup).reshape(B, 1)
    return hidden, torch.zeros(B, 1, dtype=torch.float32, device=inf.device), value, pol


class MlpInference:
    """Same surface as `SmacInference` (N, A, H, device, fused, recurrent_fused, recurrent, prediction, refresh),
    so `_DevicePlan` runs the whole search of an MLP-family model on the device as well."""

    def __init__(self, state_dict, num_agents, action_space_size, hidden_state_size, reward_support=None,
                 value_support=None, device="cuda", mode="fp32"):
        self.N, self.A, self.H = int(num_agents), int(action_space_size), int(hidden_state_size)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("MlpInference needs a CUDA device (no CPU fallback)")
        if mode not in ("fp32", "torch"):
            raise RuntimeError("MlpInference computes in fp32 (mode 'fp32' = CUDA kernel, 'torch' = torch ops on the device)")
        self.mode = mode
        self._params = {}
        sd = state_dict
        self.dyn = _Net(sd, "dynamics_network.fc_dynamic.", self.device, self._params)
        self.rew = _Net(sd, "dynamics_network.fc_reward.", self.device, self._params)
        self.val = _Net(sd, "prediction_network.fc_value.", self.device, self._params)
        self.pol = _Net(sd, "prediction_network.fc_policy.", self.device, self._params)
        D = self.N * self.H
        if self.dyn.out_features != D or self.pol.out_features != self.A or D + self.N * self.A > MAXW or self.A > 255:
            raise RuntimeError("MlpInference: network shape not supported")

        def sup(rng, size):
            # symmetric DiscreteSupport by default (config/matrix/__init__.py:22-24; core/config.py:369-381)
            lo = -(size - 1) // 2 if rng is None else int(rng[0])
            if rng is not None and int(rng[1]) - lo + 1 != size:
                raise RuntimeError("support range does not match the head's output size")
            return lo, size
        self.rsup = sup(reward_support, self.rew.out_features)
        self.vsup = sup(value_support, self.val.out_features)
        self.fused = self if mode == "fp32" else None    # `_DevicePlan` takes the one-kernel path when not None
        self.rep = RowMlp(sd, self.device, self._params) if "representation_network.mlp.0.weight" in sd else None

    @classmethod
    def from_model(cls, model, device="cuda", mode="fp32", **kw):
        return cls(model.state_dict(), int(model.num_agents), int(model.action_space_size),
                   int(model.hidden_state_size), device=device, mode=mode, **kw)

    def refresh(self, state_dict):
        for k, t in self._params.items():
            src = state_dict[k]
            if src.data_ptr() != t.data_ptr():
                t.copy_(src, non_blocking=True)
        for n in (self.dyn, self.rew, self.val, self.pol) + ((self.rep,) if self.rep is not None else ()):
            n.retranspose()

    def initial_inference(self, observation):
        """observation (B, N, ...) -> NetworkOutput-like tuple of CUDA tensors (see `initial_inference_device`)."""
        if self.rep is None:
            raise RuntimeError("the state dict holds no representation network")
        from .synthetic import NetworkOutput
        return NetworkOutput(*initial_inference_device(self, self.rep, observation))

    # ---- one launch: gather parent hidden, recurrent_inference, inverse transforms, softmax / beta ------------
    def mlp_desc(self, B, pool=None, idx_x=None, actions=None, next_hidden=None, reward=None, value=None, probs=None, beta=None,
                 greedy=None, logits_out=None, tree_agents=None, cur=-1, inv_tau=1.0):
        ptr = lambda t: (t.data_ptr() if t is not None else None)
        d = MlpDesc()
        d.B, d.N, d.A, d.H = int(B), self.N, self.A, self.H
        d.Nt = self.N if tree_agents is None else int(tree_agents)
        d.cur, d.inv_tau = int(cur), float(inv_tau)
        d.pool, d.idx_x, d.actions, d.next_hidden = ptr(pool), ptr(idx_x), ptr(actions), ptr(next_hidden)
        d.reward, d.value, d.probs, d.beta = ptr(reward), ptr(value), ptr(probs), ptr(beta)
        d.greedy, d.logits_out = ptr(greedy), ptr(logits_out)
        self.dyn.fill(d.dyn); self.rew.fill(d.rew); self.val.fill(d.val); self.pol.fill(d.pol)
        d.reward_support_min, d.reward_support_size = self.rsup
        d.value_support_min, d.value_support_size = self.vsup
        return d

    def recurrent_fused(self, B, pool, idx_x, actions, next_hidden, reward, value, probs, beta, greedy=None,
                        logits_out=None, tree_agents=None, cur=-1, inv_tau=1.0, stream=None, dbg_clock=None):
        d = self.mlp_desc(B, pool, idx_x, actions, next_hidden, reward, value, probs, beta, greedy, logits_out, tree_agents, cur,
                          inv_tau)
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        check(lib.maz_mlp_recurrent(C.byref(d), C.c_void_p(s.cuda_stream)))

    # ---- the same maths in torch ops (roots, cross-check) --------------------------------------------------------
    @staticmethod
    def _inv_transform(x, sup):
        lo, size = sup
        if size == 1:
            return x.reshape(-1)
        eps = 0.001
        s = torch.arange(lo, lo + size, dtype=torch.float32, device=x.device)
        x = (torch.softmax(x, dim=-1) * s).sum(dim=-1)
        out = ((torch.sqrt(1 + 4 * eps * (torch.abs(x) + 1 + eps)) - 1) / (2 * eps)) ** 2 - 1
        out = torch.where(x < 0, -out, out)
        out = torch.where(torch.isnan(out), torch.zeros_like(out), out)
        return torch.where(torch.abs(out) < eps, torch.zeros_like(out), out)

    @torch.no_grad()
    def prediction(self, hidden):
        B = hidden.shape[0]
        return self.pol(hidden.reshape(B * self.N, self.H)).view(B, self.N, self.A), self.val(hidden)

    @torch.no_grad()
    def recurrent(self, hidden, action, out_hidden=None):
        B = hidden.shape[0]
        onehot = F.one_hot(action.long(), num_classes=self.A).to(hidden.dtype).view(B, self.N * self.A)
        nxt = self.dyn(torch.cat([hidden, onehot], dim=1)) + hidden
        reward = self._inv_transform(self.rew(torch.cat([nxt, onehot], dim=1)), self.rsup)
        value = self._inv_transform(self.val(nxt), self.vsup)
        logits = self.pol(nxt.reshape(B * self.N, self.H)).view(B, self.N, self.A)
        if out_hidden is not None:
            out_hidden.copy_(nxt)
            nxt = out_hidden
        return nxt, reward, value, logits
