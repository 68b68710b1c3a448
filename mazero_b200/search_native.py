"""ctypes binding of the whole-search C ABI (include/maz_search.h): `NativeSearch` owns the tree arena, the hidden-state
pool and every buffer of the simulation loop, and runs one `SampledMCTS.batch_search` (core/mcts/tree_search/
mcts_sampled.py:51-200) per call -- root preparation, tree construction, all simulations, the 13 readouts -- with no
Python, torch op or host synchronisation in between.  This is the product path of `mazero_b200.mcts_sampled`; a host in
another language binds the same six functions.
"""
import ctypes as C

import numpy as np

from . import cytree
from ._lib import check, lib
from .fused import InferDesc
from .inference_mlp import MlpDesc

NET_SMAC, NET_MLP = 0, 1
AUTO, PERSISTENT, GRAPH = 0, 1, 2
STRATEGY = {"auto": AUTO, "persistent": PERSISTENT, "graph": GRAPH}

_READOUT_FIELDS = ("value", "marginal_visit_count", "marginal_priors", "num_children", "actions", "visit_count") + cytree._FLOAT_FIELDS


class SearchConfig(C.Structure):
    _fields_ = [("B", C.c_int), ("N", C.c_int), ("A", C.c_int), ("K", C.c_int), ("S", C.c_int), ("joint", C.c_int),
                ("hidden", C.c_int), ("device", C.c_int), ("strategy", C.c_int), ("net_kind", C.c_int),
                ("smac_tc", C.POINTER(InferDesc)), ("smac_small", C.POINTER(InferDesc)), ("mlp", C.POINTER(MlpDesc)),
                ("pool", C.c_void_p)]


class SearchReadout(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("values", "marginal_visit_count", "marginal_priors", "num_children", "actions",
                                          "visit_count", "pred_probs", "beta", "beta_hat", "priors", "imp_ratio", "pred_values",
                                          "mcts_values", "rewards", "qvalues")]


class SearchCall(C.Structure):
    _fields_ = [("cur", C.c_int), ("seed", C.c_uint), ("root_index_offset", C.c_uint), ("noise_eps", C.c_float), ("tau", C.c_float),
                ("pb_c_base", C.c_float), ("pb_c_init", C.c_float), ("discount", C.c_float), ("delta_lb", C.c_float),
                ("rho", C.c_float), ("lam", C.c_float),
                ("root_hidden", C.c_void_p), ("rewards", C.c_void_p), ("values", C.c_void_p), ("logits", C.c_void_p),
                ("legal", C.c_void_p), ("noise", C.c_void_p), ("factor", C.c_void_p), ("out", SearchReadout)]


class _DevArray:
    """A borrowed device buffer as a __cuda_array_interface__ object (torch.as_tensor wraps it without copying)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _fill_readout(ro, out, host):
    """out: dict name -> tensor / ndarray (keys of cytree readouts; 'value' -> values)."""
    addr = (lambda x: x.ctypes.data) if host else (lambda x: x.data_ptr())
    for k in _READOUT_FIELDS:
        x = out.get(k)
        setattr(ro, "values" if k == "value" else k, None if x is None else addr(x))


class NativeSearch:
    def __init__(self, inf, B, K, S, joint, pool=None, strategy="auto"):
        """inf: SmacInference(mode='bf16') or MlpInference(mode='fp32'); pool: optional torch CUDA tensor (S+1, B, N*H)."""
        self.inf, self.B, self.K, self.S, self.joint = inf, int(B), int(K), int(S), bool(joint)
        self.N, self.A, self.H = inf.N, inf.A, inf.H
        self.Nt = self.N if joint else 1
        cfg = SearchConfig()
        cfg.B, cfg.N, cfg.A, cfg.K, cfg.S = self.B, self.N, self.A, self.K, self.S
        cfg.joint, cfg.hidden, cfg.device = int(self.joint), self.H, inf.device.index or 0
        cfg.strategy = STRATEGY[strategy] if isinstance(strategy, str) else int(strategy)
        self._keep = []
        if hasattr(inf, "mlp_desc"):
            cfg.net_kind = NET_MLP
            d = inf.mlp_desc(self.B)
            self._keep.append(d)
            cfg.mlp = C.pointer(d)
        else:
            if getattr(inf, "fused", None) is None:
                raise RuntimeError("NativeSearch needs the fused bf16 inference (SmacInference(mode='bf16')); no fallback")
            cfg.net_kind = NET_SMAC
            tc, sm = inf.fused.weights_desc(self.B, small=False), inf.fused.weights_desc(self.B, small=True)
            self._keep += [tc, sm]
            cfg.smac_tc, cfg.smac_small = C.pointer(tc), C.pointer(sm)
        self._pool_tensor = pool
        cfg.pool = pool.data_ptr() if pool is not None else None
        self._h = C.c_void_p()
        rc = lib.maz_search_create(C.byref(self._h), C.byref(cfg))
        if rc:
            self._h = C.c_void_p()
            check(rc)
        self.strategy = {PERSISTENT: "persistent", GRAPH: "graph"}[lib.maz_search_strategy(self._h)]
        self.tree = cytree.Tree_batch.from_handle(lib.maz_search_tree(self._h), self.B, self.Nt, self.A, self.K, self.S, owner=self)
        self._call = SearchCall()
        self._calls = {}
        self._rec = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and lib is not None:
            try:
                lib.maz_search_destroy(h)
            except Exception:
                pass
            self._h = C.c_void_p()

    close = __del__

    def device_bytes(self):
        return int(lib.maz_search_device_bytes(self._h))

    def set_stream(self, cuda_stream_ptr):
        check(lib.maz_search_set_stream(self._h, C.c_void_p(int(cuda_stream_ptr) if cuda_stream_ptr else 0)))

    def check(self):
        check(lib.maz_search_check(self._h))

    def pool_ptr(self):
        return int(lib.maz_search_pool(self._h))

    def root_arrays(self):
        """(probs, beta, noises): the arrays `Tree_batch.prepare` received in the last search, as torch views (B,Nt,A)."""
        import torch

        p, b, n = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(lib.maz_search_root_arrays(self._h, C.byref(p), C.byref(b), C.byref(n)))
        shape = (self.B, self.Nt, self.A)
        return tuple(torch.as_tensor(_DevArray(x.value, shape, "<f4"), device=self.inf.device) for x in (p, b, n))

    def set_record(self, rec):
        """rec: None, or dict of CUDA tensors rewards / values (S,B), probs / beta (S,B,Nt,A), idx_x (S,B), actions (S,B,Nt)."""
        self._rec = rec
        if rec is None:
            check(lib.maz_search_set_record(self._h, None, None, None, None, None, None))
        else:
            check(lib.maz_search_set_record(self._h, *[C.c_void_p(rec[k].data_ptr()) for k in
                                                       ("rewards", "values", "probs", "beta", "idx_x", "actions")]))

    def set_debug_clock(self, t):
        check(lib.maz_search_set_debug_clock(self._h, None if t is None else C.c_void_p(t.data_ptr())))

    def set_timing(self, on=True):
        check(lib.maz_search_set_timing(self._h, int(bool(on))))

    def loop_ms(self):
        """Duration of the last search's simulation loop (persistent kernel / graph launch), CUDA events inside the library."""
        ms = C.c_float()
        check(lib.maz_search_loop_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def roots_per_cta(self):
        return int(lib.maz_search_roots_per_cta(self._h))

    def _fill(self, cur, seed, cfg, noise_eps, tau, root_index_offset):
        c = self._call
        c.cur = -1 if cur is None else int(cur)
        c.seed, c.root_index_offset = int(seed) & 0xFFFFFFFF, int(root_index_offset) & 0xFFFFFFFF
        c.noise_eps, c.tau = float(noise_eps), float(tau)
        c.pb_c_base, c.pb_c_init, c.discount = float(cfg.pb_c_base), float(cfg.pb_c_init), float(cfg.discount)
        c.delta_lb, c.rho, c.lam = float(cfg.tree_value_stat_delta_lb), float(cfg.mcts_rho), float(cfg.mcts_lambda)
        return c

    def run_dev(self, cur, seed, cfg, noise_eps, tau, rewards, values, logits, legal, noise, factor, out, root_hidden=None,
                root_index_offset=0, cache_key=None):
        """Everything device-resident (torch CUDA tensors); asynchronous on the stream set with `set_stream`.
        cache_key: when the caller passes the SAME tensors under the same key every time (a plan's own staging buffers), the
        marshalled call structure is kept and only the per-search scalars are refreshed."""
        c = self._calls.get(cache_key) if cache_key is not None else None
        if c is None:
            c = SearchCall()
            ptr = lambda t: None if t is None else t.data_ptr()
            c.root_hidden, c.rewards, c.values, c.logits = ptr(root_hidden), ptr(rewards), ptr(values), ptr(logits)
            c.legal, c.noise, c.factor = ptr(legal), ptr(noise), ptr(factor)
            _fill_readout(c.out, out, host=False)
            if cache_key is not None:
                self._calls[cache_key] = c
        self._call = c
        self._fill(cur, seed, cfg, noise_eps, tau, root_index_offset)
        check(lib.maz_search_run_dev(self._h, C.byref(c)))

    def run_host(self, cur, seed, cfg, noise_eps, tau, root_hidden, rewards, values, logits, legal, noise, factor=None,
                 root_index_offset=0):
        """Host numpy arrays in, dict of host numpy readouts out; synchronous (maz_search_run)."""
        B, N, A, K, Nt = self.B, self.N, self.A, self.K, self.Nt
        f32 = lambda x, n: None if x is None else np.ascontiguousarray(np.asarray(x, dtype=np.float32).reshape(n))
        ins = [f32(root_hidden, B * N * self.H), f32(rewards, B), f32(values, B), f32(logits, B * N * A), f32(legal, B * N * A),
               f32(noise, B * Nt * A), None if factor is None else np.ascontiguousarray(np.asarray(factor, dtype=np.int32).reshape(B * N))]
        out = {"value": np.empty(B, np.float32), "marginal_visit_count": np.empty((B, Nt, A), np.int32),
               "marginal_priors": np.empty((B, Nt, A), np.float32), "num_children": np.empty(B, np.int32),
               "actions": np.empty((B, K, Nt), np.int32), "visit_count": np.empty((B, K), np.int32)}
        out.update({k: np.empty((B, K), np.float32) for k in cytree._FLOAT_FIELDS})
        self._call = SearchCall()          # (never a cached device-pointer call)
        c = self._fill(cur, seed, cfg, noise_eps, tau, root_index_offset)
        addr = lambda x: None if x is None else x.ctypes.data
        c.root_hidden, c.rewards, c.values, c.logits, c.legal, c.noise, c.factor = [addr(x) for x in ins]
        _fill_readout(c.out, out, host=True)
        check(lib.maz_search_run(self._h, C.byref(c)))
        return out
