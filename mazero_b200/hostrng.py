"""`np_random.dirichlet([alpha] * A, rows).astype(np.float32)` (core/mcts/tree_search/mcts_sampled.py:68) with the
SAME values and the SAME final generator state, evaluated on all host cores (include/maz_hostrng.h).

Only numpy's legacy MT19937 `RandomState` with 0 < alpha < 1 is restated (the reference's workers pass such a
RandomState and alpha = 0.3, core/config.py:84); anything else is drawn by numpy itself.

`speculate` / the `Speculation` cache: right after a search has taken everything it needs from `np_random`, the draw of
the NEXT search (same shape) is started on a background thread from a copy of the generator state, while the GPU runs
the current search.  The next call uses it only if the generator is still in exactly that state (nobody drew from it in
between); otherwise it is discarded and the draw happens as usual.  Values and generator state are identical either way.
"""
import ctypes as C
import threading

import numpy as np

from ._lib import check, lib

THREADS = 0   # 0 = all hardware threads (at most 16)


def _random_state(np_random):
    if isinstance(np_random, np.random.RandomState):
        return np_random
    if np_random is np.random:
        return np.random.mtrand._rand
    return None


def _draw(key, pos, alpha, A, rows):
    """key (uint32[624], modified in place), pos -> (float32 (rows, A), new pos)"""
    p = C.c_int(int(pos))
    out = np.empty((rows, A), dtype=np.float32)
    check(lib.maz_legacy_dirichlet(C.c_void_p(key.ctypes.data), C.byref(p), float(alpha), int(rows), int(A),
                                   C.c_void_p(out.ctypes.data), None, THREADS))
    return out, p.value


def _eligible(rs, alpha, A, rows):
    return rs is not None and 0.0 < alpha < 1.0 and rows * A >= 512


class _Worker:
    """ONE persistent background thread for the speculative draws (starting a thread per search costs ~50 us)."""

    def __init__(self):
        self.cv = threading.Condition()
        self.job = None
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def _loop(self):
        while True:
            with self.cv:
                while self.job is None:
                    self.cv.wait()
                job, self.job = self.job, None
            job._run()

    def submit(self, job):
        with self.cv:
            self.job = job          # (a job that was never picked up is simply superseded: its result would be stale)
            self.cv.notify()


_worker = None


class Speculation:
    __slots__ = ("params", "key0", "pos0", "key1", "pos1", "out", "done", "error")

    def __init__(self, key0, pos0, params):
        global _worker
        self.params, self.key0, self.pos0 = params, key0, pos0
        self.key1, self.pos1, self.out, self.error = key0.copy(), pos0, None, None
        self.done = threading.Event()
        if _worker is None:
            _worker = _Worker()
        _worker.submit(self)

    def _run(self):
        try:
            alpha, A, rows = self.params
            self.out, self.pos1 = _draw(self.key1, self.pos0, alpha, A, rows)
        except Exception as e:     # the caller falls back to the ordinary draw
            self.error = e
        self.done.set()

    def wait(self, timeout=0.05):
        """True when the result is there; a job that was superseded before it started never finishes."""
        return self.done.wait(timeout)


_pending = {}   # id(RandomState) -> Speculation
STATS = {"used": 0, "discarded": 0}


def speculate(np_random, alpha, A, rows):
    """Start the next draw of this shape from the generator's CURRENT state (call it after the last draw of a search)."""
    rs = _random_state(np_random)
    alpha = float(alpha)
    if not _eligible(rs, alpha, A, rows):
        return
    st = rs.get_state(legacy=True)
    if st[0] != "MT19937":
        return
    _pending[id(rs)] = Speculation(np.array(st[1], dtype=np.uint32, order="C", copy=True), int(st[2]), (alpha, int(A), int(rows)))


def dirichlet_f32(np_random, alpha, A, rows):
    rs = _random_state(np_random)
    alpha = float(alpha)
    if _eligible(rs, alpha, A, rows):
        st = rs.get_state(legacy=True)
        if st[0] == "MT19937":
            sp = _pending.pop(id(rs), None)
            if sp is not None:
                if sp.params == (alpha, int(A), int(rows)) and sp.pos0 == int(st[2]) and np.array_equal(sp.key0, st[1]) \
                        and sp.wait() and sp.error is None:
                    rs.set_state(("MT19937", sp.key1, sp.pos1, st[3], st[4]))
                    STATS["used"] += 1
                    return sp.out
                STATS["discarded"] += 1
            key = np.array(st[1], dtype=np.uint32, order="C", copy=True)
            out, pos = _draw(key, st[2], alpha, A, rows)
            rs.set_state(("MT19937", key, pos, st[3], st[4]))
            return out
    return np_random.dirichlet([alpha] * A, rows).astype(np.float32)
