"""Clip sharding across the GPUs of one box.

Clips are independent, so the only exchange on the path is ONE all-gather of per-clip results.
Rank r owns rows [r * ceil(N/G), min(N, (r+1) * ceil(N/G))); every rank pads its block to
ceil(N/G) rows so the gather is a single `all_gather_into_tensor`, and the padding is trimmed
afterwards.  Weights, DFT/mel/DCT bases and the noise bank are replicated (each rank builds its
own Engine).  The reference has no multi-GPU path (make_submission.py runs one Keras session);
this module is what a `torchrun`-launched make_submission would call instead of
`model.predict` (make_submission.py:120-146) or the pseudo-label / vote scripts
(create_pseudo_with_thresh.py:14-43, majority_vote.py:26-56).

The collective runs on whatever backend the process group has: NCCL over NVLink on the GPU box,
gloo in the CPU tests (tests/test_shard_cpu.py) -- the shard arithmetic is backend-independent.
"""
from __future__ import annotations

import numpy as np


def rows_per_rank(n: int, world: int) -> int:
    return (n + world - 1) // world if n > 0 else 0


def shard_range(n: int, world: int, rank: int) -> tuple[int, int]:
    """[start, stop) of rank's rows; empty for trailing ranks when n < world * ceil(n/world)."""
    per = rows_per_rank(n, world)
    start = min(n, rank * per)
    return start, min(n, start + per)


def _world_rank(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def all_gather_rows(local, n: int, group=None):
    """local: torch tensor [rows_of_this_rank, ...] (any dtype / device the backend supports).
    Returns the [n, ...] concatenation over ranks in rank order (padding trimmed)."""
    import torch
    import torch.distributed as dist
    world, rank = _world_rank(group)
    if world == 1:
        return local[:n]
    per = rows_per_rank(n, world)
    s, e = shard_range(n, world, rank)
    if local.shape[0] != e - s:
        raise ValueError(f"rank {rank}: expected {e - s} local rows, got {local.shape[0]}")
    tail = tuple(local.shape[1:])
    if local.shape[0] == per and local.is_contiguous():
        padded = local
    else:
        padded = torch.zeros((per,) + tail, dtype=local.dtype, device=local.device)
        padded[: e - s] = local
    out = torch.empty((world * per,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return out[:n]


class ShardedPredictor:
    """Data-parallel wrapper: `predict(clips_of_the_whole_job)` on every rank returns the whole
    job's (probs [N,C] f32, argmax [N] int32) on every rank.

    `compute(local_clips) -> (probs, argmax)` is the rank-local device path; by default it is
    `Engine.predict_host` with the given TTA views (the CUDA path -- there is no CPU fallback).
    """

    def __init__(self, engine=None, views=((0, 1.0),), slot: int = 0, compute=None, device=None, group=None):
        self.engine, self.views, self.slot, self.group = engine, tuple(views), slot, group
        self.device = device
        if compute is None:
            if engine is None:
                raise ValueError("ShardedPredictor needs an Engine (the CUDA path) or an explicit compute callable")
            compute = lambda x: engine.predict_host(x, views=self.views, slot=self.slot)    # noqa: E731
        self.compute = compute

    def predict(self, clips: np.ndarray):
        import torch
        n = len(clips)
        world, rank = _world_rank(self.group)
        s, e = shard_range(n, world, rank)
        probs, amax = self.compute(clips[s:e])
        dev = self.device
        if dev is None:
            dev = f"cuda:{self.engine.device}" if self.engine is not None else "cpu"
        p_t = torch.as_tensor(np.ascontiguousarray(probs, np.float32)).to(dev)
        # ONE collective: the labels are the first-index argmax of the gathered mean probabilities
        # (make_submission.py:146), which is exactly what the rank-local kernel computed from the same floats
        p_all = all_gather_rows(p_t, n, self.group).cpu().numpy()
        return p_all, p_all.argmax(axis=1).astype(np.int32)


def sharded_pseudo_labels(engine, probs32_local, n: int, thresh: float, order="heng", group=None):
    """BASELINE config 4 on shards: 32->12 map + re-softmax + uint8 quantise + threshold select run
    per clip on the owning rank (kws_convert_classes / kws_select), and only the reduced
    uint8 [rows,12] / int32 label / uint8 keep blocks are gathered.
    probs32_local: torch CUDA f32 [rows_of_this_rank, 32]."""
    from .classes import class_map_32_to_12
    _, u8 = engine.convert_classes(probs32_local, class_map_32_to_12(order), 12)
    label, keep = engine.select(u8, thresh)
    return (all_gather_rows(u8, n, group), all_gather_rows(label, n, group), all_gather_rows(keep, n, group))


def sharded_vote(engine, labels_local, n: int, min_count: int = 2, group=None):
    """BASELINE config 5 on shards: labels_local torch CUDA int32 [M, rows_of_this_rank] -> voted
    labels and clear-majority flags of the whole job on every rank."""
    voted, clear = engine.vote(labels_local.contiguous(), min_count)
    return all_gather_rows(voted, n, group), all_gather_rows(clear, n, group)
