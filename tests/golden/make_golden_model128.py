"""Generate tests/golden/model128_*.npz: outputs of the REAL reference `MAMuZeroNet` at the production size
(hidden 128 per agent, config/smac/__init__.py:15-27) for the four BASELINE team shapes.

    python tests/golden/make_golden_model128.py          (needs /root/reference; ray/gymnasium are stubbed)

The network has 0.45 M parameters, too many to commit per shape, so the weights are NOT stored: they come from
`mazero_b200.synthetic.exact_state_dict(N, A, seed)` (integer-derived, bit-identical on every machine), are loaded
into the reference module here, and the fixture stores their sha256 digest next to the seeded inputs and the
reference's own outputs of `prediction(hidden)` (config/smac/model.py:494-501) and `recurrent_inference(hidden, action)`
(:562-574, inverse transforms core/config.py:430-499) in eval mode, CPU fp32.  Two chained recurrent steps are recorded
so that the second one runs on hidden states the network produced itself.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from make_golden_model import install_stubs  # noqa: E402

SHAPES = {"3m": (3, 9, 40), "2s3z": (5, 11, 24), "mmm2": (10, 18, 12), "27m": (27, 36, 5)}


def reference_model128(n, a, sd):
    from core.config import BaseConfig, DiscreteSupport
    from config.smac.model import MAMuZeroNet

    class Cfg(BaseConfig):
        def set_game(self, *a, **k): pass
        def new_game(self, *a, **k): pass
        def get_uniform_network(self): pass
        def visit_softmax_temperature_fn(self, *a, **k): pass
        def sampled_action_times_fn(self, *a, **k): pass

    cfg = object.__new__(Cfg)
    cfg.use_vectorization = True
    cfg.value_support = cfg.reward_support = DiscreteSupport(-5, 5)
    torch.manual_seed(0)
    # exactly config/smac/__init__.py:30-50 (observation size is irrelevant on the search path)
    m = MAMuZeroNet(n, (1, 16, 1, 1), a, 128, [128, 128], [128, 128], [32], [32], [32], 11, 11,
                    cfg.inverse_value_transform, cfg.inverse_reward_transform,
                    proj_hid=128, proj_out=128, pred_hid=64, pred_out=128, use_feature_norm=True)
    own = m.state_dict()
    missing = [k for k in own if k.startswith(("dynamics_network.", "prediction_network.")) and k not in sd]
    assert not missing, missing
    for k, v in sd.items():
        assert own[k].shape == v.shape, (k, own[k].shape, v.shape)
    # the positional table is a buffer computed by the reference itself: keep the reference's
    sd = {k: v for k, v in sd.items() if not k.endswith("pos_table")}
    m.load_state_dict(sd, strict=False)
    return m.eval()


def main():
    install_stubs()
    from mazero_b200.synthetic import exact_state_dict, state_dict_digest

    for name, (n, a, b) in SHAPES.items():
        seed = 100 + n
        sd = exact_state_dict(n, a, seed=seed)
        m = reference_model128(n, a, sd)
        rng = np.random.RandomState(7 + n)
        hidden = torch.from_numpy((rng.randint(-32768, 32768, size=(b, n * 128)) / 16384.0).astype(np.float32))
        action = torch.from_numpy(rng.randint(0, a, size=(b, n)).astype(np.int64))
        action2 = torch.from_numpy(rng.randint(0, a, size=(b, n)).astype(np.int64))
        with torch.no_grad():
            pol, vlog = m.prediction(hidden)
            out = m.recurrent_inference(hidden, action)
            out2 = m.recurrent_inference(out.hidden_state, action2)
        pos = m.state_dict()["dynamics_network.attention_stack.2.pos_embed.pos_table"].numpy()
        assert np.array_equal(pos, sd["dynamics_network.attention_stack.2.pos_embed.pos_table"].numpy()), "positional table differs"
        path = os.path.join(HERE, f"model128_{name}.npz")
        f32 = lambda x: np.asarray(x, dtype=np.float32)
        np.savez_compressed(
            path, dims=np.array([n, a, 128, b, seed], dtype=np.int32), digest=np.array(state_dict_digest(sd)),
            hidden=hidden.numpy(), action=action.numpy().astype(np.int32), action2=action2.numpy().astype(np.int32),
            pred_policy_logits=pol.numpy(), pred_value_logits=vlog.numpy(),
            rec_hidden=out.hidden_state.numpy(), rec_reward=f32(out.reward), rec_value=f32(out.value),
            rec_policy_logits=f32(out.policy_logits),
            rec2_hidden=out2.hidden_state.numpy(), rec2_reward=f32(out2.reward), rec2_value=f32(out2.value),
            rec2_policy_logits=f32(out2.policy_logits))
        print(name, f"{os.path.getsize(path) / 1024:.0f} KB", "reward", f32(out.reward).ravel()[:3], "value", f32(out.value).ravel()[:3],
              "logit range", float(np.abs(f32(out.policy_logits)).max()), "hidden range", float(out.hidden_state.abs().max()))


if __name__ == "__main__":
    main()
