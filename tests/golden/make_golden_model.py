"""Generate tests/golden/model_*.npz by importing the REAL reference network in this container.

    python tests/golden/make_golden_model.py          (needs /root/reference; ray/gymnasium are stubbed)

Each fixture holds the search-path subset of a reference `MAMuZeroNet.state_dict()` (random init, seed 0,
small sizes so the file stays small), seeded inputs, and the reference's own outputs of
`prediction(hidden)` and `recurrent_inference(hidden, action)` in eval mode on CPU fp32.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MAZ_REFERENCE", "/root/reference")


def install_stubs():
    ray = types.ModuleType("ray")
    ray.remote = lambda *a, **k: (a[0] if a and callable(a[0]) and not k else (lambda f: f))
    ray.get = ray.put = lambda x: x
    ray.actor = types.ModuleType("ray.actor")
    ray.actor.ActorHandle = type("ActorHandle", (), {})
    sys.modules["ray"], sys.modules["ray.actor"] = ray, ray.actor
    gym = types.ModuleType("gymnasium")
    gym.Wrapper = gym.ObservationWrapper = gym.Env = object
    gym.utils = types.ModuleType("gymnasium.utils")
    gym.utils.seeding = types.ModuleType("gymnasium.utils.seeding")
    gym.utils.seeding.np_random = lambda seed=None: (np.random.RandomState(seed), seed)
    gym.spaces = types.ModuleType("gymnasium.spaces")
    for name, mod in (("gymnasium", gym), ("gymnasium.utils", gym.utils), ("gymnasium.utils.seeding", gym.utils.seeding),
                      ("gymnasium.spaces", gym.spaces)):
        sys.modules[name] = mod
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        sys.modules.setdefault(name, types.ModuleType(name))
    cfg = types.ModuleType("config")
    cfg.__path__ = [os.path.join(REF, "config")]
    smac = types.ModuleType("config.smac")
    smac.__path__ = [os.path.join(REF, "config", "smac")]
    sys.modules["config"], sys.modules["config.smac"] = cfg, smac
    sys.path.insert(0, REF)


def reference_model(n, a, h, obs=16):
    from core.config import BaseConfig, DiscreteSupport
    from config.smac.model import MAMuZeroNet

    class Cfg(BaseConfig):
        def set_game(self, *a, **k): pass
        def new_game(self, *a, **k): pass
        def get_uniform_network(self): pass
        def visit_softmax_temperature_fn(self, *a, **k): pass
        def sampled_action_times_fn(self, *a, **k): pass

    try:
        cfg = Cfg.__new__(Cfg)
    except TypeError:
        cfg = object.__new__(Cfg)
    cfg.use_vectorization = True
    cfg.value_support = cfg.reward_support = DiscreteSupport(-5, 5)
    torch.manual_seed(0)
    m = MAMuZeroNet(n, (1, obs, 1, 1), a, h, [h, h], [h, h], [32], [32], [32], 11, 11,
                    cfg.inverse_value_transform, cfg.inverse_reward_transform,
                    proj_hid=h, proj_out=h, pred_hid=64, pred_out=h, use_feature_norm=True)
    # the constructor leaves the heads at ~0 (uniform policies); perturb so the fixture exercises them
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.05)
    return m.eval()


def reference_matrix_model(vectorized, n=2, a=3, h=64, obs=4):
    """The reference's matrix-game network exactly as config/matrix/__init__.py:11-45 builds it."""
    from core.config import BaseConfig, DiscreteSupport
    from config.matrix.model import MAMuZeroNet

    class Cfg(BaseConfig):
        def set_game(self, *a, **k): pass
        def new_game(self, *a, **k): pass
        def get_uniform_network(self): pass
        def visit_softmax_temperature_fn(self, *a, **k): pass
        def sampled_action_times_fn(self, *a, **k): pass

    cfg = object.__new__(Cfg)
    cfg.use_vectorization = vectorized
    cfg.value_support = DiscreteSupport(-10, 10) if vectorized else DiscreteSupport(0, 0)
    cfg.reward_support = DiscreteSupport(-3, 3) if vectorized else DiscreteSupport(0, 0)
    torch.manual_seed(0)
    m = MAMuZeroNet(n, (1, obs), a, h, [64, 64], [64, 64], [32], [32], [32], cfg.reward_support.size,
                    cfg.value_support.size, cfg.inverse_value_transform, cfg.inverse_reward_transform,
                    proj_hid=128, proj_out=128, pred_hid=64, pred_out=128, use_feature_norm=False)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.05)
    return m.eval(), cfg


def main_matrix():
    for name, vec in {"matrix_vec": True, "matrix_scalar": False}.items():
        m, cfg = reference_matrix_model(vec)
        n, a, h, b = 2, 3, 64, 8
        g = torch.Generator().manual_seed(2)
        hidden = torch.randn(b, n * h, generator=g)
        action = torch.randint(0, a, (b, n), generator=g)
        obs = torch.randn(b, n, 4, generator=g)
        with torch.no_grad():
            pol, vlog = m.prediction(hidden)
            out = m.recurrent_inference(hidden, action)
            ini = m.initial_inference(obs)
        keep = {k: v.numpy() for k, v in m.state_dict().items()
                if k.startswith(("dynamics_network.", "prediction_network.", "representation_network."))}
        path = os.path.join(HERE, f"model_{name}.npz")
        np.savez_compressed(
            path, dims=np.array([n, a, h, b], dtype=np.int32), hidden=hidden.numpy(), action=action.numpy().astype(np.int32),
            obs=obs.numpy(), init_hidden=ini.hidden_state.numpy(), init_value=np.asarray(ini.value, dtype=np.float32),
            init_policy_logits=np.asarray(ini.policy_logits, dtype=np.float32),
            supports=np.array([cfg.reward_support.min, cfg.reward_support.max, cfg.value_support.min, cfg.value_support.max],
                              dtype=np.int32),
            pred_policy_logits=pol.numpy(), pred_value_logits=vlog.numpy(),
            rec_hidden=out.hidden_state.numpy(), rec_reward=np.asarray(out.reward, dtype=np.float32),
            rec_value=np.asarray(out.value, dtype=np.float32), rec_policy_logits=np.asarray(out.policy_logits, dtype=np.float32),
            **{"sd." + k: v for k, v in keep.items()})
        print(name, f"{os.path.getsize(path) / 1024:.0f} KB", "value", np.asarray(out.value).ravel()[:3])


def main():
    install_stubs()
    main_matrix()
    for name, (n, a, h, b) in {"3m_h32": (3, 9, 32, 6), "2s3z_h64": (5, 11, 64, 4)}.items():
        m = reference_model(n, a, h)
        g = torch.Generator().manual_seed(2)
        hidden = torch.randn(b, n * h, generator=g)
        action = torch.randint(0, a, (b, n), generator=g)
        obs = torch.randn(b, n, 16, 1, 1, generator=g)
        with torch.no_grad():
            pol, vlog = m.prediction(hidden)
            out = m.recurrent_inference(hidden, action)
            ini = m.initial_inference(obs)
        keep = {k: v.numpy() for k, v in m.state_dict().items()
                if k.startswith(("dynamics_network.", "prediction_network.", "representation_network."))}
        path = os.path.join(HERE, f"model_{name}.npz")
        np.savez_compressed(
            path, dims=np.array([n, a, h, b], dtype=np.int32), hidden=hidden.numpy(), action=action.numpy().astype(np.int32),
            obs=obs.numpy(), init_hidden=ini.hidden_state.numpy(), init_value=np.asarray(ini.value, dtype=np.float32),
            init_policy_logits=np.asarray(ini.policy_logits, dtype=np.float32),
            pred_policy_logits=pol.numpy(), pred_value_logits=vlog.numpy(),
            rec_hidden=out.hidden_state.numpy(), rec_reward=np.asarray(out.reward, dtype=np.float32),
            rec_value=np.asarray(out.value, dtype=np.float32), rec_policy_logits=np.asarray(out.policy_logits, dtype=np.float32),
            **{"sd." + k: v for k, v in keep.items()})
        print(name, f"{os.path.getsize(path) / 1024:.0f} KB", "value", np.asarray(out.value).ravel()[:3])


if __name__ == "__main__":
    main()
