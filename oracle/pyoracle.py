"""ctypes driver for the two CPU checkers -- TEST INFRASTRUCTURE, not product code.

``OracleTreeBatch(kind="port")``  drives oracle/libmazoracle.so  (plain-C restatement, maz_oracle.c)
``OracleTreeBatch(kind="reference")`` drives oracle/_ref/libmazref.so (the reference's own C++ built by
oracle/Makefile from /root/reference; may be absent on a box that never had the reference).

The class mirrors ``cytree.Tree_batch`` (reference core/mcts/ctree/ctree_sampled/cytree.pyx:7-247):
same constructor arguments, same method names, same return types, so a parity test reads
``ours.method() == oracle.method()``.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATHS = {
    "port": (os.path.join(_HERE, "libmazoracle.so"), "mazo"),
    "reference": (os.path.join(_HERE, "_ref", "libmazref.so"), "mazref"),
}
_LIBS = {}

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int)


def build(quiet=True):
    """Compile the checkers (the C restatement always; the reference shim when /root/reference exists)."""
    out = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def available(kind):
    return os.path.exists(_PATHS[kind][0])


def _load(kind):
    if kind in _LIBS:
        return _LIBS[kind]
    path, pfx = _PATHS[kind]
    if not os.path.exists(path):
        if kind == "port":
            build()
        if not os.path.exists(path):
            raise FileNotFoundError(f"oracle library missing: {path} (run `make -C oracle`)")
    lib = C.CDLL(path)

    def fn(name, restype, *argtypes):
        f = getattr(lib, f"{pfx}_{name}")
        f.restype = restype
        f.argtypes = list(argtypes)
        return f

    api = {
        "create": fn("create", C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_uint, C.c_float, C.c_float),
        "destroy": fn("destroy", None, C.c_void_p),
        "prepare": fn("prepare", C.c_int, C.c_void_p, _f32p, _f32p, _f32p, _f32p, C.c_int, C.c_float, _f32p),
        "batch_selection": fn("batch_selection", C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_float, _i32p, _i32p, _i32p),
        "batch_expansion_and_backup": fn("batch_expansion_and_backup", C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_int, _f32p, _f32p, _f32p, _f32p),
        "get_roots_values": fn("get_roots_values", C.c_int, C.c_void_p, _f32p),
        "get_roots_marginal_visit_count": fn("get_roots_marginal_visit_count", C.c_int, C.c_void_p, _i32p),
        "get_roots_marginal_priors": fn("get_roots_marginal_priors", C.c_int, C.c_void_p, _f32p),
        "get_roots_num_children": fn("get_roots_num_children", C.c_int, C.c_void_p, _i32p),
        "readout": fn("readout", C.c_int, C.c_void_p, C.c_float, C.c_int, _i32p, _i32p, *([_f32p] * 9)),
        "stats": fn("stats", C.c_int, C.c_void_p, _i32p, _i32p),
        "last_error": fn("last_error", C.c_char_p),
    }
    _LIBS[kind] = api
    return api


def _f32(a):
    a = np.asarray(a)
    if a.dtype != np.float32:
        raise ValueError(f"Buffer dtype mismatch, expected 'float' but got '{a.dtype}'")  # typed memoryview behaviour
    return np.ascontiguousarray(a.reshape(-1))


def _p(a, t):
    return a.ctypes.data_as(t)


_FLOAT_FIELDS = ("pred_probs", "beta", "beta_hat", "priors", "imp_ratio", "pred_values", "mcts_values", "rewards", "qvalues")


class OracleTreeBatch:
    def __init__(self, root_num, agent_num, action_space_size, sampled_times, simulation_num,
                 tree_value_stat_delta_lb, random_seed, rho, lam, kind="port"):
        self._api = _load(kind)
        self.kind = kind
        self.root_num, self.agent_num, self.action_space_size = int(root_num), int(agent_num), int(action_space_size)
        self.sampled_times = int(sampled_times)
        self._h = self._api["create"](root_num, agent_num, action_space_size, sampled_times, simulation_num,
                                      tree_value_stat_delta_lb, int(random_seed) & 0xFFFFFFFF, rho, lam)
        if not self._h:
            raise RuntimeError(self._api["last_error"]().decode())

    def __del__(self):
        if getattr(self, "_h", None):
            self._api["destroy"](self._h)
            self._h = None

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self._api["last_error"]().decode())

    def prepare(self, rewards, values, policy_probs, beta, sampled_times, noise_eps, noises):
        r, v, p, b, n = _f32(rewards), _f32(values), _f32(policy_probs), _f32(beta), _f32(noises)
        self._check(self._api["prepare"](self._h, _p(r, _f32p), _p(v, _f32p), _p(p, _f32p), _p(b, _f32p),
                                         int(sampled_times), float(noise_eps), _p(n, _f32p)))

    def batch_selection(self, pb_c_base, pb_c_init, discount):
        ix = np.empty(self.root_num, dtype=np.int32)
        iy = np.empty(self.root_num, dtype=np.int32)
        act = np.empty(self.root_num * self.agent_num, dtype=np.int32)
        self._check(self._api["batch_selection"](self._h, pb_c_base, pb_c_init, discount, _p(ix, _i32p), _p(iy, _i32p), _p(act, _i32p)))
        return ix.tolist(), iy.tolist(), act.reshape(self.root_num, self.agent_num)

    def batch_expansion_and_backup(self, hidden_state_index_x, discount, sampled_times, rewards, values, policy_probs, beta):
        r, v, p, b = _f32(rewards), _f32(values), _f32(policy_probs), _f32(beta)
        self._check(self._api["batch_expansion_and_backup"](self._h, int(hidden_state_index_x), float(discount), int(sampled_times),
                                                            _p(r, _f32p), _p(v, _f32p), _p(p, _f32p), _p(b, _f32p)))

    def get_roots_values(self):
        out = np.empty(self.root_num, dtype=np.float32)
        self._check(self._api["get_roots_values"](self._h, _p(out, _f32p)))
        return out

    def get_roots_marginal_visit_count(self):
        out = np.empty(self.root_num * self.agent_num * self.action_space_size, dtype=np.int32)
        self._check(self._api["get_roots_marginal_visit_count"](self._h, _p(out, _i32p)))
        return out.reshape(self.root_num, self.agent_num, self.action_space_size)

    def get_roots_marginal_priors(self):
        out = np.empty(self.root_num * self.agent_num * self.action_space_size, dtype=np.float32)
        self._check(self._api["get_roots_marginal_priors"](self._h, _p(out, _f32p)))
        return out.reshape(self.root_num, self.agent_num, self.action_space_size)

    def get_roots_num_children(self):
        out = np.empty(self.root_num, dtype=np.int32)
        self._check(self._api["get_roots_num_children"](self._h, _p(out, _i32p)))
        return out

    def readout(self, discount):
        """All 11 per-root-child readouts as padded (B,K[,N]) arrays + num_children."""
        B, K, N = self.root_num, self.sampled_times, self.agent_num
        nc = self.get_roots_num_children()
        res = {"num_children": nc,
               "actions": np.zeros((B, K, N), dtype=np.int32),
               "visit_count": np.zeros((B, K), dtype=np.int32)}
        for f in _FLOAT_FIELDS:
            res[f] = np.zeros((B, K), dtype=np.float32)
        self._check(self._api["readout"](self._h, float(discount), K, _p(res["actions"], _i32p), _p(res["visit_count"], _i32p),
                                         *[_p(res[f], _f32p) for f in _FLOAT_FIELDS]))
        return res

    def _ragged(self, name, discount=0.0):
        r = self.readout(discount)
        return [r[name][b, : r["num_children"][b]].copy() for b in range(self.root_num)]

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
        tot = np.empty(self.root_num, dtype=np.int32)
        sl = np.empty(self.root_num, dtype=np.int32)
        self._check(self._api["stats"](self._h, _p(tot, _i32p), _p(sl, _i32p)))
        return tot, sl
