"""`Tree_batch`: the B200 replacement of the reference's Cython class `cytree.Tree_batch`
(core/mcts/ctree/ctree_sampled/cytree.pyx:7-247) -- same constructor, same method names, same argument
meaning, same return types, same error behaviour -- backed by libmaz_b200.so (CUDA, sm_100a).

    from mazero_b200 import cytree
    trees = cytree.Tree_batch(B, agent_num, A, K, S, delta_lb, seed, rho, lam)
    trees.prepare(rewards, values, policy_probs, beta, K, noise_eps, noises)
    ix, iy, actions = trees.batch_selection(pb_c_base, pb_c_init, discount)
    trees.batch_expansion_and_backup(sim + 1, discount, K, rewards, values, policy_probs, beta)
    trees.get_roots_values() ...

numpy inputs take the host-pointer entry points (what the reference binding passes); torch CUDA tensors
take the `_dev` entry points: nothing is copied and nothing synchronises (see `*_device` methods).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib

_FLOAT_FIELDS = ("pred_probs", "beta", "beta_hat", "priors", "imp_ratio", "pred_values", "mcts_values", "rewards",
                 "qvalues")


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _host_f32(a, name):
    """Typed-memoryview semantics of cytree.pyx:22-45: reshape(-1), make contiguous, float32 only."""
    a = np.asarray(a).reshape(-1)
    if a.dtype != np.float32:
        raise ValueError(f"Buffer dtype mismatch, expected 'float' but got '{a.dtype.name}' ({name})")
    if not a.flags["C_CONTIGUOUS"]:
        a = np.ascontiguousarray(a)
    return a


def _dev_ptr(x, dtype_name, numel, name):
    import torch

    want = {"float32": torch.float32, "int32": torch.int32}[dtype_name]
    if not x.is_cuda:
        raise ValueError(f"{name}: expected a CUDA tensor")
    if x.dtype != want:
        raise ValueError(f"Buffer dtype mismatch, expected '{dtype_name}' but got '{x.dtype}' ({name})")
    if not x.is_contiguous():
        raise ValueError(f"{name}: device tensors must be contiguous")
    if x.numel() != numel:
        raise ValueError(f"{name}: expected {numel} elements, got {x.numel()}")
    return C.c_void_p(x.data_ptr())


class Tree_batch:
    def __init__(self, root_num, agent_num, action_space_size, sampled_times, simulation_num,
                 tree_value_stat_delta_lb, random_seed, rho, lam, device=None, root_index_offset=0):
        self.root_num = int(root_num)
        self.agent_num = int(agent_num)
        self.action_space_size = int(action_space_size)
        self.sampled_times = int(sampled_times)
        self.simulation_num = int(simulation_num)
        self._h = C.c_void_p()
        self._cache = {}
        self._cache_dev = {}     # marshalled device pointers of the last readout_device target
        seed = int(random_seed) & 0xFFFFFFFF
        if device is None:
            rc = lib.maz_tree_create(C.byref(self._h), self.root_num, self.agent_num, self.action_space_size,
                                     self.sampled_times, self.simulation_num, float(tree_value_stat_delta_lb), seed,
                                     float(rho), float(lam))
        else:
            rc = lib.maz_tree_create_ex(C.byref(self._h), self.root_num, self.agent_num, self.action_space_size,
                                        self.sampled_times, self.simulation_num, float(tree_value_stat_delta_lb), seed,
                                        float(rho), float(lam), int(device), int(root_index_offset) & 0xFFFFFFFF)
        if rc != _lib.MAZ_OK:
            self._h = C.c_void_p()
            check(rc)

    @classmethod
    def from_handle(cls, handle, root_num, agent_num, action_space_size, sampled_times, simulation_num, owner=None):
        """View of a tree batch that another native object owns (the whole-search handle of include/maz_search.h)."""
        t = cls.__new__(cls)
        t.root_num, t.agent_num, t.action_space_size = int(root_num), int(agent_num), int(action_space_size)
        t.sampled_times, t.simulation_num = int(sampled_times), int(simulation_num)
        t._h = C.c_void_p(int(handle))
        t._cache, t._cache_dev = {}, {}
        t._borrowed = True
        import weakref

        t._owner = None if owner is None else weakref.ref(owner)
        return t

    def __del__(self):
        if getattr(self, "_borrowed", False):
            return
        h = getattr(self, "_h", None)
        if h and lib is not None:   # `lib` is None while the interpreter shuts down
            try:
                lib.maz_tree_destroy(h)
            except Exception:
                pass
            self._h = C.c_void_p()

    # ---- handle management (extensions; the reference builds a new object per search) -----------------
    def reset(self, random_seed, tree_value_stat_delta_lb, rho, lam, root_index_offset=0):
        self._cache.clear()
        check(lib.maz_tree_reset(self._h, int(random_seed) & 0xFFFFFFFF, float(tree_value_stat_delta_lb), float(rho),
                                 float(lam), int(root_index_offset) & 0xFFFFFFFF))

    def set_stream(self, cuda_stream_ptr):
        check(lib.maz_tree_set_stream(self._h, C.c_void_p(int(cuda_stream_ptr) if cuda_stream_ptr else 0)))

    def set_puct(self, pb_c_base, pb_c_init):
        check(lib.maz_tree_set_puct(self._h, float(pb_c_base), float(pb_c_init)))

    def check(self):
        check(lib.maz_tree_check(self._h))

    def arena_bytes(self):
        return int(lib.maz_tree_arena_bytes(self._h))

    # ---- the four step methods (cytree.pyx:21-91) --------------------------------------------------------
    def prepare(self, rewards, values, policy_probs, beta, sampled_times, noise_eps, noises):
        self._cache.clear()
        B, NA = self.root_num, self.agent_num * self.action_space_size
        if _is_torch(rewards):
            check(lib.maz_tree_prepare_dev(
                self._h, _dev_ptr(rewards, "float32", B, "rewards"), _dev_ptr(values, "float32", B, "values"),
                _dev_ptr(policy_probs, "float32", B * NA, "policy_probs"), _dev_ptr(beta, "float32", B * NA, "beta"),
                int(sampled_times), float(noise_eps), _dev_ptr(noises, "float32", B * NA, "noises")))
            return
        r, v = _host_f32(rewards, "rewards"), _host_f32(values, "values")
        p, b, n = _host_f32(policy_probs, "policy_probs"), _host_f32(beta, "beta"), _host_f32(noises, "noises")
        for a, want, name in ((r, B, "rewards"), (v, B, "values"), (p, B * NA, "policy_probs"), (b, B * NA, "beta"),
                              (n, B * NA, "noises")):
            if a.size != want:
                raise ValueError(f"{name}: expected {want} elements, got {a.size}")
        check(lib.maz_tree_prepare(self._h, r.ctypes.data, v.ctypes.data, p.ctypes.data, b.ctypes.data,
                                   int(sampled_times), float(noise_eps), n.ctypes.data))

    def batch_selection(self, pb_c_base, pb_c_init, discount):
        B, N = self.root_num, self.agent_num
        ix = np.empty(B, dtype=np.int32)
        iy = np.empty(B, dtype=np.int32)
        act = np.empty(B * N, dtype=np.int32)
        check(lib.maz_tree_batch_selection(self._h, float(pb_c_base), float(pb_c_init), float(discount),
                                           ix.ctypes.data, iy.ctypes.data, act.ctypes.data))
        return ix.tolist(), iy.tolist(), act.reshape(B, N)

    def batch_selection_device(self, pb_c_base, pb_c_init, discount, idx_x, idx_y, actions):
        """Asynchronous selection into caller-owned int32 CUDA tensors (B,), (B,), (B,N)."""
        B, N = self.root_num, self.agent_num
        check(lib.maz_tree_batch_selection_dev(
            self._h, float(pb_c_base), float(pb_c_init), float(discount), _dev_ptr(idx_x, "int32", B, "idx_x"),
            _dev_ptr(idx_y, "int32", B, "idx_y"), _dev_ptr(actions, "int32", B * N, "actions")))

    def batch_expansion_and_backup(self, hidden_state_index_x, discount, sampled_times, rewards, values, policy_probs, beta):
        self._cache.clear()
        B, NA = self.root_num, self.agent_num * self.action_space_size
        if _is_torch(rewards):
            check(lib.maz_tree_batch_expansion_and_backup_dev(
                self._h, int(hidden_state_index_x), float(discount), int(sampled_times),
                _dev_ptr(rewards, "float32", B, "rewards"), _dev_ptr(values, "float32", B, "values"),
                _dev_ptr(policy_probs, "float32", B * NA, "policy_probs"), _dev_ptr(beta, "float32", B * NA, "beta")))
            return
        r, v = _host_f32(rewards, "rewards"), _host_f32(values, "values")
        p, b = _host_f32(policy_probs, "policy_probs"), _host_f32(beta, "beta")
        for a, want, name in ((r, B, "rewards"), (v, B, "values"), (p, B * NA, "policy_probs"), (b, B * NA, "beta")):
            if a.size != want:
                raise ValueError(f"{name}: expected {want} elements, got {a.size}")
        check(lib.maz_tree_batch_expansion_and_backup(self._h, int(hidden_state_index_x), float(discount),
                                                      int(sampled_times), r.ctypes.data, v.ctypes.data, p.ctypes.data,
                                                      b.ctypes.data))

    def expansion_backup_selection_device(self, hidden_state_index_x, discount, sampled_times, rewards, values, policy_probs,
                                          beta, pb_c_base, pb_c_init, idx_x, idx_y, actions):
        """Fused step of the on-device loop: batch_expansion_and_backup(s) then batch_selection(s+1), one launch."""
        self._cache.clear()
        B, N, NA = self.root_num, self.agent_num, self.agent_num * self.action_space_size
        check(lib.maz_tree_expansion_backup_selection_dev(
            self._h, int(hidden_state_index_x), float(discount), int(sampled_times),
            _dev_ptr(rewards, "float32", B, "rewards"), _dev_ptr(values, "float32", B, "values"),
            _dev_ptr(policy_probs, "float32", B * NA, "policy_probs"), _dev_ptr(beta, "float32", B * NA, "beta"),
            float(pb_c_base), float(pb_c_init), _dev_ptr(idx_x, "int32", B, "idx_x"), _dev_ptr(idx_y, "int32", B, "idx_y"),
            _dev_ptr(actions, "int32", B * N, "actions")))

    # ---- readouts (cytree.pyx:93-241) ------------------------------------------------------------------
    def readout(self, discount=0.0):
        """Everything observable at the roots as padded arrays (one kernel, one round of copies)."""
        key = float(discount)
        if key in self._cache:
            return self._cache[key]
        B, N, A, K = self.root_num, self.agent_num, self.action_space_size, self.sampled_times
        r = {
            "value": np.empty(B, dtype=np.float32),
            "marginal_visit_count": np.empty((B, N, A), dtype=np.int32),
            "marginal_priors": np.empty((B, N, A), dtype=np.float32),
            "num_children": np.empty(B, dtype=np.int32),
            "actions": np.empty((B, K, N), dtype=np.int32),
            "visit_count": np.empty((B, K), dtype=np.int32),
        }
        for f in _FLOAT_FIELDS:
            r[f] = np.empty((B, K), dtype=np.float32)
        check(lib.maz_tree_readout(self._h, key, r["value"].ctypes.data, r["marginal_visit_count"].ctypes.data,
                                   r["marginal_priors"].ctypes.data, r["num_children"].ctypes.data,
                                   r["actions"].ctypes.data, r["visit_count"].ctypes.data,
                                   *[r[f].ctypes.data for f in _FLOAT_FIELDS]))
        self._cache[key] = r
        return r

    def readout_device(self, discount, out):
        """Asynchronous readout into a dict of caller-owned CUDA tensors (keys as in `readout`; any may be absent)."""
        B, N, A, K = self.root_num, self.agent_num, self.action_space_size, self.sampled_times
        spec = [("value", "float32", B), ("marginal_visit_count", "int32", B * N * A), ("marginal_priors", "float32", B * N * A),
                ("num_children", "int32", B), ("actions", "int32", B * K * N), ("visit_count", "int32", B * K)]
        spec += [(f, "float32", B * K) for f in _FLOAT_FIELDS]
        # the on-device search reads out into the same dict of views every time: validate and marshal the pointers once
        ck = ("readout_dev", id(out))
        ptrs = self._cache_dev.get(ck)
        if ptrs is None or any(p.value != (out[k].data_ptr() if k in out and out[k] is not None else None)
                               for p, (k, _, _) in zip(ptrs, spec)):
            ptrs = [(_dev_ptr(out[k], dt, n, k) if k in out and out[k] is not None else C.c_void_p(0)) for k, dt, n in spec]
            self._cache_dev = {ck: ptrs}
        check(lib.maz_tree_readout_dev(self._h, float(discount), *ptrs))

    def _ragged(self, name, discount=0.0):
        r = self.readout(discount)
        nc = r["num_children"]
        return [r[name][b, : nc[b]].copy() for b in range(self.root_num)]

    def get_roots_values(self):
        return self.readout()["value"].copy()

    def get_roots_marginal_visit_count(self):
        return self.readout()["marginal_visit_count"].copy()

    def get_roots_marginal_priors(self):
        return self.readout()["marginal_priors"].copy()

    def get_roots_num_children(self):
        return self.readout()["num_children"].copy()

    def get_roots_sampled_visit_count(self):
        return self._ragged("visit_count")

    def get_roots_sampled_actions(self):
        return self._ragged("actions")

    def get_roots_sampled_pred_probs(self):
        return self._ragged("pred_probs")

    def get_roots_sampled_beta(self):
        return self._ragged("beta")

    def get_roots_sampled_beta_hat(self):
        return self._ragged("beta_hat")

    def get_roots_sampled_priors(self):
        return self._ragged("priors")

    def get_roots_sampled_imp_ratio(self):
        return self._ragged("imp_ratio")

    def get_roots_sampled_pred_values(self):
        return self._ragged("pred_values")

    def get_roots_sampled_mcts_values(self):
        return self._ragged("mcts_values")

    def get_roots_sampled_rewards(self):
        return self._ragged("rewards")

    def get_roots_sampled_qvalues(self, discount):
        return self._ragged("qvalues", discount)

    def stats(self):
        """(tot_nodes (B,), last_search_len (B,), mean search depth, mean children per expanded node)."""
        B = self.root_num
        tot = np.empty(B, dtype=np.int32)
        sl = np.empty(B, dtype=np.int32)
        s1, s2 = C.c_longlong(0), C.c_longlong(0)
        check(lib.maz_tree_stats(self._h, tot.ctypes.data, sl.ctypes.data, C.byref(s1), C.byref(s2)))
        return tot, sl, int(s1.value), int(s2.value)

    def print(self):
        """Debug dump (cytree.pyx:243-247 prints the whole pool to stderr; here: a per-tree summary)."""
        import sys

        tot, sl, s1, s2 = self.stats()
        r = self.readout()
        for b in range(self.root_num):
            print(f"---------- Tree {b} info ----------\n\tnodes: {tot[b]}, last search_len: {sl[b]}, "
                  f"root children: {r['num_children'][b]}, root value: {r['value'][b]}", file=sys.stderr)
