"""CPU: host-side weight packing of the two-tiles-in-flight inference kernel (mazero_b200/fused.py, csrc/infer_twin.cuh):
32 matrices in the order of the kernel's stage table, stored as K-halves in the tcgen05 core-matrix layout, one-hot weight
blocks as fp32 [A][128] tables of the bf16-rounded weights behind the parameter vector; descriptor fields; kernel selection."""
import torch

from mazero_b200 import fused
from mazero_b200.synthetic import random_state_dict

H, GH, PH = 128, 64, 32


def _unpack_operand(flat, rows, k):
    """inverse of fused.pack_operand (no swizzle): [rows/8][k/8][8][8] -> [rows, k]"""
    return flat.view(rows // 8, k // 8, 8, 8).permute(0, 2, 1, 3).reshape(rows, k)


def _expected_matrices(sd, A, KA):
    g = lambda k: sd[k].float()
    d, p = "dynamics_network.", "prediction_network."
    w_in = g(d + "attention_stack.0.weight")
    out = [w_in[:, :H]]
    for l in range(3):
        e = f"{d}attention_stack.2.encoder.layers.{l}."
        qkv = g(e + "self_attn.in_proj_weight")
        q, k, v = qkv[0:H], qkv[H:2 * H], qkv[2 * H:3 * H]
        for h in (0, 1):      # four heads per stage: [Q rows | K rows] stacked, then the V rows
            out += [torch.cat([q[64 * h:64 * h + 64], k[64 * h:64 * h + 64]], 0), v[64 * h:64 * h + 64]]
        out += [g(e + "self_attn.out_proj.weight"), g(e + "linear1.weight"), g(e + "linear2.weight")]
    wd1 = g(d + "fc_dynamic.0.weight")
    out += [wd1[:, :H], wd1[:, H + A:], g(d + "fc_dynamic.3.weight"), g(d + "fc_dynamic.6.weight")]
    r, v = d + "reward_predictor.", p + "value_predictor."
    wr1 = torch.cat([g(r + "gc1.lin_layer.weight"), g(r + "nn_gc1.weight")], 0)
    out += [wr1[:, :H], torch.cat([g(r + "gc2.lin_layer.weight"), g(r + "nn_gc2.weight")], 0),
            torch.cat([g(v + "gc1.lin_layer.weight"), g(v + "nn_gc1.weight")], 0),
            torch.cat([g(v + "gc2.lin_layer.weight"), g(v + "nn_gc2.weight")], 0), g(p + "fc_policy.0.weight")]
    pol3 = torch.zeros(KA, PH)
    pol3[:A] = g(p + "fc_policy.3.weight")
    out.append(pol3)
    return out, (w_in, wd1, wr1)


def test_twin_packing_roundtrip(built_lib):
    N, A = 5, 11
    sd = random_state_dict(N, A, seed=3)
    fp = fused.FusedParams(sd, N, A, "cpu")
    mats, (w_in, wd1, wr1) = _expected_matrices(sd, A, fp.KA)
    assert len(mats) == fused.NCHUNK == len(fp.chunk_bytes_t)
    raw = fp.wpk_t
    for i, m in enumerate(mats):
        rows, k = m.shape
        assert fp.chunk_off_t[i] % 16 == 0 and fp.chunk_bytes_t[i] == rows * k * 2
        base = fp.chunk_off_t[i] // 2
        got = torch.empty(rows, k, dtype=torch.bfloat16)
        off = base
        for k0 in range(0, k, 64):                       # pieces of <= 64 input features, each <= 16 KB (one ring slot)
            kk = min(64, k - k0)
            assert rows * kk * 2 <= 16384
            got[:, k0:k0 + kk] = _unpack_operand(raw[off:off + rows * kk], rows, kk)
            off += rows * kk
        assert torch.equal(got, m.to(torch.bfloat16)), f"matrix {i}"
    # one-hot blocks: table[a][col] = bf16(W[col, H + a]), 32-byte aligned behind the fp32 parameter vector
    assert fp.vec.numel() % 8 == 0 and torch.equal(fp.vec_t[:fp.vec.numel()], fp.vec)
    for off, w in zip(fp.off_oh, (w_in, wd1, wr1)):
        assert off % 8 == 0
        tab = fp.vec_t[off:off + A * H // 2].contiguous().view(torch.bfloat16).view(A, H)     # [A][64] words = [A][128] bf16
        assert torch.equal(tab, w[:, H:H + A].to(torch.bfloat16).t())
    assert fp.vec_t.numel() == fp.off_oh[2] + A * H // 2


def test_twin_descriptor_and_selection(built_lib, monkeypatch):
    N, A = 10, 18
    fp = fused.FusedParams(random_state_dict(N, A, seed=1), N, A, "cpu")
    d1 = fp.desc(8192, None, None, None, None, None, None, None, None, twin=True)
    assert d1.tc_layout == 1 and d1.vec_floats == fp.vec_t.numel() and d1.wpk == fp.wpk_t.data_ptr()
    assert (d1.o_oh_in, d1.o_oh_dyn, d1.o_oh_rg) == tuple(fp.off_oh)
    d0 = fp.desc(8192, None, None, None, None, None, None, None, None, twin=False)
    assert d0.tc_layout == 0 and d0.vec_floats == fp.vec.numel() and d0.wpk == fp.wpk.data_ptr()
    ds = fp.desc(8192, None, None, None, None, None, None, None, None, small=True, twin=True)     # the small-batch layout ignores it
    assert ds.tc_layout == 0 and ds.wpk == fp.wpk_h.data_ptr()
    # selection by the wave cost model (pairs: ~141 us per wave; single tiles: ~80 + 3.2 N us per wave); env override
    monkeypatch.delenv("MAZ_INFER_TC", raising=False)
    assert fused.use_twin(8192, 10) and not fused.use_twin(1024, 10) and fused.use_twin(64, 27)
    assert fused.use_twin(4096, 5) and not fused.use_twin(2048, 3)
    assert fused.use_twin(8192, 3) and not fused.use_twin(16384, 3)          # 2 x 88 us vs 141 us; 3 x 88 us vs 2 x 141 us
    monkeypatch.setenv("MAZ_INFER_TC", "v1")
    assert not fused.use_twin(8192, 10)
    monkeypatch.setenv("MAZ_INFER_TC", "twin")
    assert fused.use_twin(16, 3)
