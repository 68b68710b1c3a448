"""Generate tests/golden/turns_kat.npz with the reference's OWN `select_action` / `eps_greedy_action`
(core/utils.py:289-334), imported in this container (needs /root/reference; gymnasium is stubbed).

    python tests/golden/make_golden_turns.py

Each case: ragged visit counts, a temperature, a seed -> the reference's (action_pos, entropy) with
np.random.RandomState(seed), together with the one uniform that RandomState yields, so that the result can be
reproduced with an injected draw; and eps-greedy cases with torch.manual_seed(seed)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_model as mg  # noqa: E402


def reference_utils():
    mg.install_stubs()
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_core_utils", os.path.join(mg.REF, "core", "utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ru = reference_utils()
    rng = np.random.RandomState(0)
    M, K = 400, 10
    counts = np.zeros((M, K), np.int32)
    lens = rng.randint(1, K + 1, size=M).astype(np.int32)
    temps = rng.choice([1.0, 0.5, 0.25, 0.7], size=M)
    pos, ent, uni = np.zeros(M, np.int32), np.zeros(M), np.zeros(M)
    for i in range(M):
        c = rng.multinomial(50, rng.dirichlet([0.5] * lens[i])).astype(np.int32)
        if c.sum() == 0:
            c[0] = 1
        counts[i, : lens[i]] = c
        uni[i] = np.random.RandomState(1000 + i).random_sample()
        p, e = ru.select_action(c, temperature=float(temps[i]), deterministic=False, np_random=np.random.RandomState(1000 + i))
        pos[i], ent[i] = p, e
    # eps-greedy: torch global RNG, per call rand_like(scalar) then Categorical(mask).sample()
    E, A = 200, 9
    masks = (rng.rand(E, A) < 0.6).astype(np.float32)
    masks[:, 1] = 1
    greedy = rng.randint(0, A, size=E).astype(np.int32)
    eps = 0.3
    eps_u, rand_act, res = np.zeros(E, np.float32), np.zeros(E, np.int32), np.zeros(E, np.int32)
    for i in range(E):
        torch.manual_seed(i)
        res[i] = int(ru.eps_greedy_action(int(greedy[i]), masks[i], eps)[0])
        torch.manual_seed(i)
        m = torch.from_numpy(masks[i])
        eps_u[i] = float(torch.rand_like(m[..., 0].float()))
        rand_act[i] = int(torch.distributions.Categorical(m).sample())
    path = os.path.join(HERE, "turns_kat.npz")
    np.savez_compressed(path, counts=counts, lens=lens, temps=temps, uniforms=uni, pos=pos, entropy=ent,
                        masks=masks, greedy=greedy, eps=np.float32(eps), eps_u=eps_u, rand_act=rand_act, eps_result=res)
    print(path, os.path.getsize(path), "bytes; picked random in", int((eps_u < eps).sum()), "of", E)


if __name__ == "__main__":
    main()
