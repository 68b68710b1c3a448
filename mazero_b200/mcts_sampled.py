"""`SampledMCTS.batch_search` / `SearchOutput`: the search driver, API-compatible with the reference's
core/mcts/tree_search/mcts_sampled.py:12-200, so core/selfplay_worker.py:201-211 and
core/reanalyze_worker.py:121-129,284-294,540-550 can call it unchanged.

Two execution paths, same results contract:

* device path  -- the model is a SMAC `MAMuZeroNet`-shaped network (or a `SmacInference`): the whole
  select -> gather hidden -> recurrent_inference -> softmax/beta -> expand+backup loop runs on the GPU
  as one CUDA graph; the host touches nothing between `prepare` and the readout.
* step path    -- any other `BaseNet` (e.g. the reference's mock models): the reference's simulation loop
  verbatim, with `cytree.Tree_batch` (CUDA) in place of the Cython tree.  Still no CPU tree.

Extension over the fork's signature: `current_agent_idx=None` selects JOINT mode (tree agent_num =
true_num_agents, the upstream MAZero behaviour the BASELINE configs describe); an int selects the fork's
sequential-agent mode (tree agent_num = 1, mcts_sampled.py:53,89).
"""
from collections.abc import Sequence
from typing import List, NamedTuple, Optional, Tuple

import numpy as np
import torch

from . import cytree
from .inference import SmacInference
from .inference_mlp import MlpInference
from . import inference_mlp


class SearchOutput(NamedTuple):  # mcts_sampled.py:12-26
    value: np.ndarray
    marginal_visit_count: np.ndarray
    marginal_priors: np.ndarray
    sampled_actions: List[np.ndarray]
    sampled_visit_count: List[np.ndarray]
    sampled_pred_probs: List[np.ndarray]
    sampled_beta: List[np.ndarray]
    sampled_beta_hat: List[np.ndarray]
    sampled_priors: List[np.ndarray]
    sampled_imp_ratio: List[np.ndarray]
    sampled_pred_values: List[np.ndarray]
    sampled_mcts_values: List[np.ndarray]
    sampled_rewards: List[np.ndarray]
    sampled_qvalues: List[np.ndarray]


class Ragged(Sequence):
    """Lazy list-of-arrays view over a padded (B,K,...) array + per-root lengths: element b is
    padded[b, :n[b]].  The reference materialises 11*B tiny numpy arrays per search
    (cytree.pyx:111-241); callers only index `[b]`, iterate, and use `.size` / `[pos, 0]` / np.sum."""

    __slots__ = ("_pad", "_n")

    def __init__(self, padded, n):
        self._pad, self._n = padded, n

    def __len__(self):
        return len(self._n)

    def __getitem__(self, b):
        if isinstance(b, slice):
            return [self[i] for i in range(*b.indices(len(self)))]
        if b < 0:
            b += len(self)
        return self._pad[b, : self._n[b]]

    def __eq__(self, other):
        return len(self) == len(other) and all(np.array_equal(a, b) for a, b in zip(self, other))


def _softmax_np(logits):
    p = np.exp(logits - np.max(logits, axis=-1, keepdims=True))
    return p / np.sum(p, axis=-1, keepdims=True)


def _output_from_readout(r):
    n = r["num_children"]
    rg = lambda k: Ragged(r[k], n)
    return SearchOutput(r["value"], r["marginal_visit_count"], r["marginal_priors"], rg("actions"), rg("visit_count"),
                        rg("pred_probs"), rg("beta"), rg("beta_hat"), rg("priors"), rg("imp_ratio"), rg("pred_values"),
                        rg("mcts_values"), rg("rewards"), rg("qvalues"))


class _DevicePlan:
    """Static buffers + the captured CUDA graphs of S simulations for one problem shape.  The sequential-agent turns
    (current_agent_idx = 0..N-1) share ONE plan -- tree arena, hidden-state pool and staging buffers are by far the
    largest allocations -- and differ only in their captured graph (`graphs[cur]`)."""

    def __init__(self, inf: SmacInference, B, K, S, cur, cfg, tau, use_graph=True):
        self.inf, self.B, self.K, self.S, self.cur, self.tau = inf, B, K, S, cur, float(tau)
        self.N, self.A, self.H = inf.N, inf.A, inf.H
        self.Nt = self.N if cur is None else 1  # agents in the tree
        dev = inf.device
        self.dev = dev
        self.c_base, self.c_init, self.discount = float(cfg.pb_c_base), float(cfg.pb_c_init), float(cfg.discount)
        self.tree = cytree.Tree_batch(B, self.Nt, self.A, K, S, float(cfg.tree_value_stat_delta_lb), 0,
                                      float(cfg.mcts_rho), float(cfg.mcts_lambda), device=dev.index or 0)
        self.tree.set_puct(self.c_base, self.c_init)
        f32, i32 = torch.float32, torch.int32
        D = self.N * self.H
        self.pool = torch.zeros(S + 1, B, D, dtype=f32, device=dev)          # hidden_states_pool (mcts_sampled.py:86)
        self.greedy = torch.zeros(S + 1, B, self.N, dtype=i32, device=dev)   # argmax_a prediction(node).policy (A.11)
        self.idx_x = torch.zeros(B, dtype=i32, device=dev)
        self.idx_y = torch.zeros(B, dtype=i32, device=dev)
        self.act = torch.zeros(B, self.Nt, dtype=i32, device=dev)
        self.factor = torch.zeros(B, max(self.N, 1), dtype=i32, device=dev)
        self.rows = torch.arange(B, dtype=torch.int64, device=dev)
        z = lambda *s: torch.zeros(*s, dtype=f32, device=dev)
        self.root_r, self.root_v = z(B), z(B)
        self.root_p, self.root_b, self.root_n = z(B, self.Nt, self.A), z(B, self.Nt, self.A), z(B, self.Nt, self.A)
        # per-simulation network outputs (fused path writes them in place)
        self.sim_r, self.sim_v = z(B), z(B)
        self.sim_p, self.sim_b = z(B, self.Nt, self.A), z(B, self.Nt, self.A)
        self.graphs = {}   # current_agent_idx (or None) -> captured CUDA graph
        self.use_graph = use_graph
        self.record = None  # when a list: per-simulation injected arrays are appended (parity replay tests)
        # all readouts live in ONE device buffer mirrored by ONE pinned host buffer: one D2H copy per search
        spec = [("value", f32, (B,)), ("marginal_visit_count", i32, (B, self.Nt, self.A)),
                ("marginal_priors", f32, (B, self.Nt, self.A)), ("num_children", i32, (B,)),
                ("actions", i32, (B, K, self.Nt)), ("visit_count", i32, (B, K))]
        spec += [(f, f32, (B, K)) for f in cytree._FLOAT_FIELDS]
        words = sum(int(np.prod(shp)) for _, _, shp in spec)
        self.out_flat = torch.zeros(words, dtype=i32, device=dev)
        self.out_host = torch.zeros(words, dtype=i32).pin_memory()
        self.out, self.out_np = {}, {}
        o = 0
        host_np = self.out_host.numpy()
        for name, dt, shp in spec:
            n = int(np.prod(shp))
            self.out[name] = self.out_flat[o:o + n].view(dt).view(*shp)
            self.out_np[name] = host_np[o:o + n].view(np.float32 if dt is f32 else np.int32).reshape(shp)
            o += n

    def _simulate(self, s, select_first=True, select_next=False):
        """One simulation.  select_first: run the selection kernel at the start (else the previous simulation's
        fused expand+backup+select launch already did); select_next: fuse the NEXT simulation's selection into
        this simulation's expansion+backup launch."""
        t, B = self.tree, self.B
        if select_first:
            t.batch_selection_device(self.c_base, self.c_init, self.discount, self.idx_x, self.idx_y, self.act)
        flat = None
        if self.cur is None:
            joint = self.act
        else:                                                                 # mcts_sampled.py:116-147
            flat = self.idx_x.long() * B + self.rows                          # row of the parent in the pool
            g = self.greedy.view(-1, self.N).index_select(0, flat)
            joint = torch.cat([self.factor[:, : self.cur], self.act, g[:, self.cur + 1:]], dim=1).contiguous()
        if self.inf.fused is not None:
            # ONE kernel: gather parent hidden, recurrent_inference on the tensor cores, inverse support
            # transforms, softmax / beta of the tree agents, greedy actions of every agent
            self.inf.recurrent_fused(B, self.pool, self.idx_x, joint, self.pool[s + 1], self.sim_r, self.sim_v, self.sim_p,
                                     self.sim_b, self.greedy[s + 1] if self.cur is not None else None, None, self.Nt,
                                     -1 if self.cur is None else self.cur, 1.0 / self.tau)
            if self.record is not None:
                self.record.append((self.sim_r.clone(), self.sim_v.clone(), self.sim_p.clone(), self.sim_b.clone(),
                                    self.idx_x.clone(), self.act.clone()))
            if select_next:
                t.expansion_backup_selection_device(s + 1, self.discount, self.K, self.sim_r, self.sim_v, self.sim_p,
                                                    self.sim_b, self.c_base, self.c_init, self.idx_x, self.idx_y, self.act)
            else:
                t.batch_expansion_and_backup(s + 1, self.discount, self.K, self.sim_r, self.sim_v, self.sim_p, self.sim_b)
            return
        if flat is None:
            flat = self.idx_x.long() * B + self.rows
        h = self.pool.view(-1, self.N * self.H).index_select(0, flat)
        _, rew, val, logits = self.inf.recurrent(h, joint, out_hidden=self.pool[s + 1])
        if self.cur is not None:
            self.greedy[s + 1].copy_(logits.argmax(dim=-1))
            logits = logits[:, self.cur: self.cur + 1]
        p = torch.softmax(logits, dim=-1)                                     # mcts_sampled.py:158-161
        b = p if self.tau == 1.0 else p ** (1.0 / self.tau)
        b = b / b.sum(dim=-1, keepdim=True)
        p, b = p.contiguous(), b.contiguous()
        if self.record is not None:
            self.record.append((rew.clone(), val.clone(), p.clone(), b.clone(), self.idx_x.clone(), self.act.clone()))
        if select_next:
            t.expansion_backup_selection_device(s + 1, self.discount, self.K, rew, val, p, b, self.c_base, self.c_init,
                                                self.idx_x, self.idx_y, self.act)
        else:
            t.batch_expansion_and_backup(s + 1, self.discount, self.K, rew, val, p, b)

    def _loop(self):
        # two launches per simulation: [inference] -> [expand + backup + next selection]
        fuse = self.record is None
        for s in range(self.S):
            self._simulate(s, select_first=(s == 0 or not fuse), select_next=(fuse and s + 1 < self.S))

    @property
    def graph(self):
        return self.graphs.get(self.cur)

    def run(self, seed, cfg, noise_eps, root_hidden, rewards, values, probs, beta, noises, root_greedy, factor,
            root_index_offset=0, cur="same"):
        """root_* host numpy (pinned or not) or device tensors; returns the padded readout dict (host numpy)."""
        if cur != "same":
            assert (cur is None) == (self.cur is None), "a plan is either joint or sequential"
            self.cur = cur
        stream = torch.cuda.current_stream(self.dev)
        self.tree.set_stream(stream.cuda_stream)
        cp = lambda dst, src: dst.copy_(src if torch.is_tensor(src) else torch.from_numpy(np.ascontiguousarray(src)), non_blocking=True)
        cp(self.pool[0], root_hidden.reshape(self.B, -1))
        cp(self.root_r, rewards); cp(self.root_v, values)
        cp(self.root_p, probs); cp(self.root_b, beta); cp(self.root_n, noises)
        if self.cur is not None:
            cp(self.greedy[0], root_greedy)
            if factor is not None and self.cur > 0:
                cp(self.factor[:, : self.cur], factor[:, : self.cur] if torch.is_tensor(factor)
                   else np.asarray(factor, dtype=np.int32)[:, : self.cur])

        def prepare():
            self.tree.reset(seed, float(cfg.tree_value_stat_delta_lb), float(cfg.mcts_rho), float(cfg.mcts_lambda), root_index_offset)
            self.tree.prepare(self.root_r, self.root_v, self.root_p, self.root_b, self.K, float(noise_eps), self.root_n)

        if self.use_graph and self.record is None:
            if self.cur not in self.graphs:
                prepare()          # the capture warm-up runs one real simulation: it needs valid roots
                self._capture()
            prepare()
            self.graphs[self.cur].replay()
        else:
            prepare()
            self._loop()
        self.tree.readout_device(self.discount, self.out)
        self.out_host.copy_(self.out_flat, non_blocking=True)
        self.tree.check()          # synchronises the stream and surfaces device-side invariant failures
        return {k: v.copy() for k, v in self.out_np.items()}

    def _capture(self):
        cur = torch.cuda.current_stream(self.dev)
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):   # warm-up outside capture (cuBLAS handles, workspaces, allocator)
            self.tree.set_stream(side.cuda_stream)
            self._simulate(0)
            self.tree.check()
        cur.wait_stream(side)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.tree.set_stream(torch.cuda.current_stream(self.dev).cuda_stream)
            self._loop()
        self.tree.set_stream(cur.cuda_stream)
        self.graphs[self.cur] = g


class SampledMCTS(object):
    def __init__(self, config, np_random: np.random.RandomState = None, use_cuda_graph: bool = True,
                 inference_mode: str = "auto"):
        """inference_mode: "bf16" = fused tensor-core kernel, "fp32" = parity mode (plain fp32 torch ops on the
        device), "auto" = bf16 when the network has the reference SMAC architecture, else fp32."""
        self.config = config
        self.np_random = np.random if np_random is None else np_random
        self.use_cuda_graph = use_cuda_graph
        self.inference_mode = inference_mode
        self._inference = {}
        self._plans = {}

    # ---- model adaptation ---------------------------------------------------------------------------------
    def _device_inference(self, model, device) -> Optional[SmacInference]:
        if isinstance(model, (SmacInference, MlpInference)):
            return model
        if not isinstance(model, torch.nn.Module):
            return None
        try:
            sd = model.state_dict()
        except Exception:
            return None
        smac = "dynamics_network.attention_stack.0.weight" in sd and "prediction_network.fc_policy.0.weight" in sd
        mlp = not smac and inference_mlp.supported(sd)     # the matrix-game family (config/matrix/model.py)
        if not smac and not mlp:
            return None
        dev = torch.device(device) if device is not None else sd["prediction_network.fc_policy.0.weight"].device
        if dev.type != "cuda":
            return None
        key = (id(model), str(dev))
        inf = self._inference.get(key)
        version = sum(int(getattr(p, "_version", 0)) for p in sd.values())
        if inf is None and mlp:
            inf = MlpInference.from_model(model, device=dev, mode="torch" if self.inference_mode == "torch" else "fp32")
            inf._src_version = version
            self._inference[key] = inf
        elif inf is None:
            mode = self.inference_mode
            if mode == "auto":
                from . import fused

                h = int(getattr(model, "hidden_state_size_per_agent", getattr(model, "hidden", 128)))
                mode = "bf16" if fused.supported(sd, int(model.num_agents), int(model.action_space_size), h) else "fp32"
            inf = SmacInference.from_model(model, device=dev, mode=mode)
            inf._src_version = version
            self._inference[key] = inf
        elif getattr(inf, "_src_version", None) != version:   # weights were updated in place (set_weights)
            inf.refresh(sd)
            inf._src_version = version
        return inf

    # ---- the entry point (mcts_sampled.py:34-46) -----------------------------------------------------------
    def batch_search(
        self,
        model,
        network_output,
        current_agent_idx: Optional[int],
        factor: np.ndarray,
        true_num_agents: int,
        legal_actions_lst: np.ndarray = None,
        device: torch.device = None,
        add_noise: bool = False,
        sampled_tau: float = 1.0,
        sampled_actions_res: Tuple[np.ndarray, np.ndarray] = None,
        root_index_offset: int = 0,
    ) -> SearchOutput:
        cfg = self.config
        noise_alpha, noise_epsilon = cfg.root_dirichlet_alpha, cfg.root_exploration_fraction
        A, K = cfg.action_space_size, cfg.sampled_action_times
        joint = current_agent_idx is None
        Nt = true_num_agents if joint else 1
        B = network_output.hidden_state.shape[0]

        def host(x):
            return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)

        batch_rewards, batch_values = host(network_output.reward), host(network_output.value)
        all_logits = host(network_output.policy_logits)
        if joint:
            logits = all_logits.reshape(B, Nt, A)
        else:
            logits = all_logits[:, current_agent_idx, :].reshape(B, 1, A)           # mcts_sampled.py:60-61
        assert batch_values.shape == (B, 1) and logits.shape == (B, Nt, A)
        probs = _softmax_np(logits)                                                  # :64-65

        # exploration noise (:68-70): drawn even when add_noise is False, one Dirichlet per (root, tree agent)
        noises = self.np_random.dirichlet([noise_alpha] * A, B * Nt if joint else B).astype(np.float32).reshape(B, Nt, A)
        if not add_noise:
            noise_epsilon = 0.0
        legal = None
        if legal_actions_lst is not None:                                            # :73-83
            legal = (legal_actions_lst.reshape(B, Nt, A) if joint
                     else legal_actions_lst[:, current_agent_idx, :].reshape(B, 1, A))
            probs *= legal
            probs += legal * 1e-4
            assert ~(np.sum(probs, axis=-1) == 0).sum()
            probs = probs / np.sum(probs, axis=-1, keepdims=True)
            noises *= legal
            noises += legal * 1e-4
            noises = noises / np.sum(noises, axis=-1, keepdims=True)

        seed = self.np_random.choice(256)                                            # :89 (after the Dirichlet)
        if sampled_actions_res is not None:
            raise NotImplementedError                                                # :108-109
        beta = probs * (1 - noise_epsilon) + noises * noise_epsilon                  # :93-100
        beta = beta ** (1 / sampled_tau)
        if legal is not None:
            beta *= legal
            assert ~(np.sum(beta, axis=-1) == 0).sum()
        beta = beta / np.sum(beta, axis=-1, keepdims=True)
        batch_rewards = batch_rewards.reshape(B).astype(np.float32)
        batch_values = batch_values.reshape(B).astype(np.float32)
        probs, beta, noises = probs.astype(np.float32), beta.astype(np.float32), noises.astype(np.float32)

        inf = self._device_inference(model, device)
        if inf is not None:
            key = (id(inf), B, K, cfg.num_simulations, joint, float(sampled_tau))   # all sequential turns share a plan
            plan = self._plans.get(key)
            if plan is None:
                plan = _DevicePlan(inf, B, K, cfg.num_simulations, current_agent_idx, cfg, sampled_tau, self.use_cuda_graph)
                self._plans[key] = plan
            root_greedy = np.argmax(all_logits.reshape(B, true_num_agents, A), axis=-1).astype(np.int32)
            r = plan.run(int(seed), cfg, noise_epsilon, network_output.hidden_state, batch_rewards, batch_values, probs, beta,
                         noises, root_greedy, factor, root_index_offset, cur=current_agent_idx)
            return _output_from_readout(r)
        return self._search_step_path(model, network_output, current_agent_idx, factor, true_num_agents, device, sampled_tau,
                                      seed, noise_epsilon, batch_rewards, batch_values, probs, beta, noises, root_index_offset)

    # ---- reference simulation loop with the CUDA tree (any BaseNet) --------------------------------------------
    def _search_step_path(self, model, network_output, cur, factor, true_num_agents, device, sampled_tau, seed,
                          noise_epsilon, batch_rewards, batch_values, probs, beta, noises, root_index_offset):
        cfg = self.config
        A, K = cfg.action_space_size, cfg.sampled_action_times
        joint = cur is None
        Nt = true_num_agents if joint else 1
        B = network_output.hidden_state.shape[0]
        hidden_states_pool = [network_output.hidden_state]
        dev_index = None
        if device is not None and torch.device(device).type == "cuda":
            dev_index = torch.device(device).index or 0
        trees = cytree.Tree_batch(B, Nt, A, K, cfg.num_simulations, cfg.tree_value_stat_delta_lb, seed, cfg.mcts_rho,
                                  cfg.mcts_lambda, device=dev_index, root_index_offset=root_index_offset)
        trees.prepare(batch_rewards, batch_values, probs, beta, K, noise_epsilon, noises)
        with torch.no_grad():
            if hasattr(model, "eval"):
                model.eval()
            for index_simulation in range(cfg.num_simulations):
                ix, iy, batch_actions = trees.batch_selection(cfg.pb_c_base, cfg.pb_c_init, cfg.discount)
                hs = torch.vstack([hidden_states_pool[x][y] for x, y in zip(ix, iy)])
                if joint:
                    joint_action = batch_actions.astype(np.int32)
                else:
                    joint_action = np.zeros((B, true_num_agents), dtype=np.int32)
                    if factor is not None:
                        for k in range(cur):
                            joint_action[:, k] = factor[:, k]
                    joint_action[:, cur] = batch_actions.squeeze()
                    if cur + 1 < true_num_agents:
                        prior_logits, _ = model.prediction(hs)
                        if isinstance(prior_logits, torch.Tensor):
                            prior_logits = prior_logits.cpu().numpy()
                        for k in range(cur + 1, true_num_agents):
                            joint_action[:, k] = np.argmax(prior_logits[:, k, :], axis=-1)
                out = model.recurrent_inference(hs, torch.from_numpy(joint_action).to(device))
                hidden_states_pool.append(out.hidden_state)
                host = lambda x: x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)
                pl = host(out.policy_logits)
                pl = pl.reshape(B, Nt, A) if joint else pl[:, cur, :].reshape(B, 1, A)
                p = _softmax_np(pl)
                b = p ** (1 / sampled_tau)
                b = b / np.sum(b, axis=-1, keepdims=True)
                trees.batch_expansion_and_backup(index_simulation + 1, cfg.discount, K,
                                                 host(out.reward).reshape(B).astype(np.float32),
                                                 host(out.value).reshape(B).astype(np.float32),
                                                 p.astype(np.float32), b.astype(np.float32))
        return _output_from_readout(trees.readout(cfg.discount))
