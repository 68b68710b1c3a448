"""CPU (gloo, world_size 2): host-side logic of the multi-GPU path -- root slices, global-root-index offsets,
and the single all-gather of packed readouts per search (mazero_b200/sharding.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mazero_b200.sharding import pack_readout, shard_range, unpack_readout


def test_shard_ranges_partition_the_roots():
    for total in (1, 7, 1024, 1025, 4099):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                s, c = shard_range(total, r, world)
                seen += list(range(s, s + c))
            assert seen == list(range(total))


def test_pack_unpack_roundtrip():
    rng = np.random.RandomState(0)
    ro = {"value": rng.randn(5).astype(np.float32), "visit_count": rng.randint(0, 9, (5, 10)).astype(np.int32),
          "actions": rng.randint(0, 9, (5, 10, 3)).astype(np.int32), "qvalues": rng.randn(5, 10).astype(np.float32)}
    mat, layout = pack_readout(ro)
    back = unpack_readout(mat, layout)
    for k in ro:
        assert back[k].dtype == ro[k].dtype and np.array_equal(back[k], ro[k])


def _fake_readout(start, count):
    """Deterministic per-root 'search results' keyed by the GLOBAL root index."""
    g = np.arange(start, start + count)
    return {"value": (g * 0.5).astype(np.float32), "num_children": (g % 7).astype(np.int32),
            "visit_count": (g[:, None] * 10 + np.arange(4)[None]).astype(np.int32)}


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mazero_b200.sharding import all_gather_readouts

    start, count = shard_range(total, rank, world)
    out = all_gather_readouts(_fake_readout(start, count), total)
    q.put((rank, {k: v.tolist() for k, v in out.items()}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 11])
def test_all_gather_readouts_gloo_world2(total):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _fake_readout(0, total)
    for rank, out in results:
        for k in want:
            assert np.array_equal(np.asarray(out[k], dtype=want[k].dtype), want[k]), (rank, k)


# ---- the product path of `SampledMCTS.batch_search_sharded`: field-major packed blocks, padded to the largest shard, ONE
# all_gather_into_tensor, merged in global root order (host-side logic; the device search itself is covered on the GPU) -------
_SPEC = [("value", np.float32, None), ("num_children", np.int32, None), ("visit_count", np.int32, (4,))]


def _spec(rows):
    return [(n, dt, (rows,) + (shp or ())) for n, dt, shp in _SPEC]


def _packed_block(start, count, rows):
    from mazero_b200.sharding import pad_rows

    ro = _fake_readout(start, count)
    words = [np.ascontiguousarray(pad_rows(ro[n], rows)).reshape(-1).view(np.int32) for n, _, _ in _SPEC]
    return np.concatenate(words)


def _worker_packed(rank, world, port, total, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mazero_b200.sharding import merge_shards

    start, count = shard_range(total, rank, world)
    rows = shard_range(total, 0, world)[1]
    mine = torch.from_numpy(_packed_block(start, count, rows))
    gathered = torch.empty(world * mine.numel(), dtype=torch.int32)
    dist.all_gather_into_tensor(gathered, mine)
    out = merge_shards(gathered.numpy().reshape(world, -1), _spec(rows), total, world)
    q.put((rank, {k: v.tolist() for k, v in out.items()}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 11])
def test_packed_blocks_merge_in_global_root_order_gloo_world2(total):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_packed, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _fake_readout(0, total)
    for rank, out in results:
        for k in want:
            assert np.array_equal(np.asarray(out[k], dtype=want[k].dtype), want[k]), (rank, k)


def test_pad_rows_repeats_the_last_root():
    from mazero_b200.sharding import pad_rows

    x = np.arange(6, dtype=np.float32).reshape(3, 2)
    y = pad_rows(x, 5)
    assert y.shape == (5, 2) and np.array_equal(y[:3], x) and np.array_equal(y[3], x[2]) and np.array_equal(y[4], x[2])
    t = pad_rows(torch.from_numpy(x), 4)
    assert tuple(t.shape) == (4, 2) and torch.equal(t[3], t[2])
    assert pad_rows(x, 3) is x
