"""CPU: the production-size (hidden 128) network against outputs of the REAL reference MAMuZeroNet
(tests/golden/model128_*.npz; config/smac/model.py:494-501,562-574; core/config.py:430-499):
  * the oracle restatement (oracle/model_oracle.py), and
  * the product's fp32 parity mode `SmacInference(mode="fp32")` (plain torch ops; the floating-point reference the
    CUDA kernels are compared with on the GPU) -- so that "parity with the fp32 twin" is parity with the reference."""
import numpy as np
import pytest
import torch

from _golden128 import GOLDEN128, IDS128, load128

TOL = dict(rtol=1e-5, atol=2e-6)      # same ops, fp32 round-off only
TOL_SCALAR = dict(rtol=1e-4, atol=1e-5)   # oracle: the reference's own op sequence behind softmax . support and inv_h
# product twin: same formula, different op order.  inv_h (core/config.py:430-442) computes sqrt(1 + 4e-3 (|x|+1.001)) - 1 ~ 2e-3 in
# fp32: the cancellation turns one ulp of the square root (6e-8) into 3e-5 relative, i.e. ~1e-4 absolute on the scalar.
TOL_SCALAR_TWIN = dict(rtol=1e-4, atol=3e-4)


def test_fixtures_exist():
    assert len(GOLDEN128) == 4


@pytest.mark.parametrize("path", GOLDEN128, ids=IDS128)
def test_oracle_model_h128_matches_reference(path):
    from oracle.model_oracle import OracleMAMuZeroNet

    z, sd, (n, a, h, b) = load128(path)
    m = OracleMAMuZeroNet(n, a, hidden_state_size=h).load_reference_state_dict(sd).eval()
    hidden = torch.from_numpy(z["hidden"])
    with torch.no_grad():
        pol, vlog = m.prediction(hidden)
        nxt, rew, val, plog = m.recurrent_inference(hidden, torch.from_numpy(z["action"]))
        nxt2, rew2, val2, plog2 = m.recurrent_inference(nxt, torch.from_numpy(z["action2"]))
    np.testing.assert_allclose(pol.numpy(), z["pred_policy_logits"], **TOL)
    np.testing.assert_allclose(vlog.numpy(), z["pred_value_logits"], **TOL)
    for got, key, tol in ((nxt, "rec_hidden", TOL), (plog, "rec_policy_logits", TOL), (rew, "rec_reward", TOL_SCALAR),
                          (val, "rec_value", TOL_SCALAR), (nxt2, "rec2_hidden", TOL), (plog2, "rec2_policy_logits", TOL),
                          (rew2, "rec2_reward", TOL_SCALAR), (val2, "rec2_value", TOL_SCALAR)):
        np.testing.assert_allclose(got.numpy().reshape(z[key].shape), z[key], err_msg=key, **tol)


@pytest.mark.parametrize("path", GOLDEN128, ids=IDS128)
def test_fp32_parity_mode_matches_reference_on_cpu(path):
    from mazero_b200.inference import SmacInference

    z, sd, (n, a, h, b) = load128(path)
    inf = SmacInference(sd, n, a, device="cpu", mode="fp32")
    hidden = torch.from_numpy(z["hidden"])
    pol, vlog = inf.prediction(hidden)
    nxt, rew, val, plog = inf.recurrent(hidden, torch.from_numpy(z["action"]))
    nxt2, rew2, val2, plog2 = inf.recurrent(nxt, torch.from_numpy(z["action2"]))
    np.testing.assert_allclose(pol.numpy(), z["pred_policy_logits"], **TOL)
    np.testing.assert_allclose(vlog.numpy(), z["pred_value_logits"], **TOL)
    for got, key, tol in ((nxt, "rec_hidden", TOL), (plog, "rec_policy_logits", TOL), (rew, "rec_reward", TOL_SCALAR_TWIN),
                          (val, "rec_value", TOL_SCALAR_TWIN), (nxt2, "rec2_hidden", TOL), (plog2, "rec2_policy_logits", TOL),
                          (rew2, "rec2_reward", TOL_SCALAR_TWIN), (val2, "rec2_value", TOL_SCALAR_TWIN)):
        np.testing.assert_allclose(got.numpy().reshape(z[key].shape), z[key], err_msg=key, **tol)
