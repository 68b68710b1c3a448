"""Generate tests/golden/tree_*.npz from the COMPILED REFERENCE (oracle/_ref/libmazref.so).

Run in the build container (where /root/reference exists):
    make -C oracle && python tests/golden/make_golden.py

Each fixture stores the injected float32 inputs themselves (numpy's exp/softmax is not bit-stable across
CPUs, so regenerating them from a seed on another box would not be safe) and every observable output of
the reference tree: per-simulation selection results and all readouts.
The reference's own tests pin shapes only (unit_test_mcts.py:203-210); these are the known-answer
vectors this repo pins numerics with.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

from oracle.pyoracle import OracleTreeBatch  # noqa: E402
from _harness import MCTS, Inputs, drive, pack  # noqa: E402

# name: (B, N, A, K, S, mode, legal_frac, tree_seed, rho)
CASES = {
    "matrix_joint": (4, 2, 3, 5, 50, "random", None, 11, 0.75),
    "matrix_seq_mock": (4, 1, 3, 5, 50, "mock", 0.7, 200, 0.75),
    "3m_joint": (4, 3, 9, 10, 50, "random", None, 77, 0.75),
    "3m_seq_legal": (4, 1, 9, 10, 50, "random", 0.7, 5, 0.25),
    "3m_quantized": (4, 3, 9, 10, 30, "quantized", None, 255, 0.75),
    "2s3z_joint": (2, 5, 11, 10, 100, "random", None, 31, 0.75),
    "mmm2_joint": (2, 10, 18, 10, 50, "random", 0.7, 3, 0.75),
    "27m_joint": (1, 27, 36, 10, 30, "random", None, 9, 0.75),
    "k1_chain": (4, 2, 4, 1, 40, "random", None, 1, 0.75),
    "a1_nodraw": (3, 2, 1, 3, 10, "random", None, 8, 0.75),
}


def main():
    for name, (B, N, A, K, S, mode, legal, seed, rho) in CASES.items():
        inp = Inputs(B, N, A, S, seed=seed, mode=mode, legal_frac=legal)
        mcts = dict(MCTS, rho=rho)
        ref = OracleTreeBatch(B, N, A, K, S, mcts["delta_lb"], seed, rho, mcts["lam"], kind="reference")
        out = pack(drive(ref, inp, K, mcts=mcts))
        tot, _ = ref.stats()
        path = os.path.join(HERE, f"tree_{name}.npz")
        np.savez_compressed(
            path,
            dims=np.array([B, N, A, K, S], dtype=np.int32), tree_seed=np.uint32(seed),
            mcts=np.array([mcts["pb_c_base"], mcts["pb_c_init"], mcts["discount"], rho, mcts["lam"], mcts["delta_lb"]], dtype=np.float32),
            in_rewards=inp.rewards, in_values=inp.values, in_probs=inp.probs, in_beta=inp.beta, in_noises=inp.noises,
            in_noise_eps=inp.noise_eps, tot_nodes=tot, **{"out_" + k: v for k, v in out.items()})
        print(f"{name}: wrote {os.path.getsize(path) / 1024:.1f} KB, root values {out['value'][:2]}")


if __name__ == "__main__":
    main()
