"""`np_random.dirichlet([alpha] * A, rows).astype(np.float32)` (core/mcts/tree_search/mcts_sampled.py:68) with the
SAME values and the SAME final generator state, evaluated on all host cores (include/maz_hostrng.h).

Only numpy's legacy MT19937 `RandomState` with 0 < alpha < 1 is restated (the reference's workers pass such a
RandomState and alpha = 0.3, core/config.py:84); anything else is drawn by numpy itself."""
import ctypes as C

import numpy as np

from ._lib import check, lib

THREADS = 0   # 0 = all hardware threads (at most 16)


def _random_state(np_random):
    if isinstance(np_random, np.random.RandomState):
        return np_random
    if np_random is np.random:
        return np.random.mtrand._rand
    return None


def dirichlet_f32(np_random, alpha, A, rows):
    rs = _random_state(np_random)
    alpha = float(alpha)
    if rs is not None and 0.0 < alpha < 1.0 and rows * A >= 512:
        st = rs.get_state(legacy=True)
        if st[0] == "MT19937":
            key = np.array(st[1], dtype=np.uint32, order="C", copy=True)
            pos = C.c_int(int(st[2]))
            out = np.empty((rows, A), dtype=np.float32)
            check(lib.maz_legacy_dirichlet(C.c_void_p(key.ctypes.data), C.byref(pos), alpha, int(rows), int(A),
                                           C.c_void_p(out.ctypes.data), None, THREADS))
            rs.set_state(("MT19937", key, pos.value, st[3], st[4]))
            return out
    return np_random.dirichlet([alpha] * A, rows).astype(np.float32)
