"""Loader of tests/golden/model128_*.npz: outputs of the REAL reference MAMuZeroNet at hidden 128 (made by
tests/golden/make_golden_model128.py); the weights are regenerated from the recorded seed and checked by digest."""
import glob
import os

import numpy as np

GOLDEN128 = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "model128_*.npz")))
IDS128 = [os.path.basename(p)[9:-4] for p in GOLDEN128]


def load128(path):
    from mazero_b200.synthetic import exact_state_dict, state_dict_digest

    z = np.load(path)
    n, a, h, b, seed = [int(x) for x in z["dims"]]
    sd = exact_state_dict(n, a, seed=seed)
    assert state_dict_digest(sd) == str(z["digest"]), "regenerated weights differ from the ones the fixture was made with"
    return z, sd, (n, a, h, b)
