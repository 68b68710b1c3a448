"""CPU: the torch-fp32 restatement of the network (oracle/model_oracle.py) against golden outputs of the
REAL reference MAMuZeroNet (tests/golden/model_*.npz, made by tests/golden/make_golden_model.py)."""
import glob
import os

import numpy as np
import pytest
import torch

MODEL_GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "model_*.npz")))


def load_model_golden(path):
    from oracle.model_oracle import OracleMAMuZeroNet, OracleMatrixMuZeroNet

    z = np.load(path)
    n, a, h, b = [int(x) for x in z["dims"]]
    sd = {k[3:]: z[k] for k in z.files if k.startswith("sd.")}
    obs = int(np.prod(z["obs"].shape[2:])) if "obs" in z.files else None
    if "supports" in z.files:     # the matrix-game network (config/matrix/model.py)
        r0, r1, v0, v1 = [int(x) for x in z["supports"]]
        m = OracleMatrixMuZeroNet(n, a, h, reward_support=(r0, r1), value_support=(v0, v1), obs_size=obs)
        return z, m.load_reference_state_dict(sd).eval(), sd, (n, a, h, b)
    m = OracleMAMuZeroNet(n, a, hidden_state_size=h, fc_dynamic_layers=(h, h), obs_size=obs).load_reference_state_dict(sd).eval()
    return z, m, sd, (n, a, h, b)


@pytest.mark.parametrize("path", MODEL_GOLDEN, ids=[os.path.basename(p)[6:-4] for p in MODEL_GOLDEN])
def test_oracle_model_matches_reference_outputs(path):
    z, m, _, _ = load_model_golden(path)
    hidden, action = torch.from_numpy(z["hidden"]), torch.from_numpy(z["action"])
    with torch.no_grad():
        pol, vlog = m.prediction(hidden)
        nxt, rew, val, plog = m.recurrent_inference(hidden, action)
    tol = dict(rtol=1e-5, atol=1e-6)  # same ops, same library, same box: fp32 round-off only
    np.testing.assert_allclose(pol.numpy(), z["pred_policy_logits"], **tol)
    np.testing.assert_allclose(vlog.numpy(), z["pred_value_logits"], **tol)
    np.testing.assert_allclose(nxt.numpy(), z["rec_hidden"], **tol)
    np.testing.assert_allclose(plog.numpy(), z["rec_policy_logits"], **tol)
    np.testing.assert_allclose(rew.numpy(), z["rec_reward"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(val.numpy(), z["rec_value"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("path", MODEL_GOLDEN, ids=[os.path.basename(p)[6:-4] for p in MODEL_GOLDEN])
def test_oracle_initial_inference_matches_reference_outputs(path):
    """representation network + prediction + inverse value transform (config/smac/model.py:542-559, config/matrix/model.py:345-356)"""
    z, m, _, _ = load_model_golden(path)
    hidden, value, pol = m.initial_inference(torch.from_numpy(z["obs"]))
    tol = dict(rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(hidden.numpy(), z["init_hidden"], **tol)
    np.testing.assert_allclose(pol.numpy(), z["init_policy_logits"], **tol)
    np.testing.assert_allclose(value.numpy().reshape(-1), z["init_value"].reshape(-1), rtol=1e-4, atol=1e-5)
