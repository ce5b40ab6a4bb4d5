"""N > 1 host path on CPU: shard arithmetic and the single all-gather, world_size 2 and 3 over
gloo (the GPU box runs the same code over NCCL).  The rank-local compute is a deterministic
stand-in here (the product's compute is CUDA-only); what is tested is partition, padding,
gather order and trimming -- including the reference job size 158,538 = 8 * 19,817 + 2
(convert_from_see_v3_bugfix.py:66)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from speech_recognition_b200 import sharded


def test_shard_ranges_cover_job_exactly():
    n = 158538
    for world in (1, 2, 4, 8):
        per = sharded.rows_per_rank(n, world)
        spans = [sharded.shard_range(n, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert all(e - s <= per for s, e in spans)
    assert sharded.rows_per_rank(n, 8) == 19818
    assert sharded.shard_range(n, 8, 7) == (7 * 19818, n)          # last rank is 6 rows short
    # fewer clips than ranks: trailing ranks are empty, nothing is lost
    assert [sharded.shard_range(3, 8, r) for r in range(8)] == [(0, 1), (1, 2), (2, 3)] + [(3, 3)] * 5
    assert sharded.shard_range(0, 4, 2) == (0, 0)


def _fake_compute(clips):
    """Deterministic per-clip 'probabilities' that depend only on the clip's content."""
    key = clips[:, 0].astype(np.float64)
    probs = np.stack([np.sin(key * (c + 1)) ** 2 for c in range(12)], 1).astype(np.float32)
    probs /= probs.sum(1, keepdims=True) + 1e-9
    return probs, probs.argmax(1).astype(np.int32)


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.RandomState(7)
        clips = rng.randn(n, 4).astype(np.float32)                 # same job on every rank
        pred = sharded.ShardedPredictor(compute=_fake_compute, device="cpu")
        probs, amax = pred.predict(clips)
        ref_p, ref_a = _fake_compute(clips)
        ok = np.array_equal(probs, ref_p) and np.array_equal(amax, ref_a) and probs.shape == (n, 12)
        # uint8 / multi-dim payloads (config 4 gathers uint8 [rows,12] + int32 labels)
        s, e = sharded.shard_range(n, world, rank)
        u8 = torch.arange(n * 12, dtype=torch.int64).reshape(n, 12).remainder(251).to(torch.uint8)
        got = sharded.all_gather_rows(u8[s:e].clone(), n)
        ok = ok and torch.equal(got, u8)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,n", [(2, 101), (2, 64), (3, 2), (2, 0)])
def test_gather_over_gloo(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(r, True) for r in range(world)]


def test_single_process_is_identity():
    x = torch.arange(10, dtype=torch.float32).reshape(5, 2)
    assert torch.equal(sharded.all_gather_rows(x, 5), x)
