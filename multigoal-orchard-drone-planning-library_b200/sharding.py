"""Batch sharding for one-process-per-GPU runs.

Every state / motion / goal sample is independent given the (replicated, read-only) scene, so a batch is split into
contiguous index ranges whose boundaries are multiples of 32: result bitmask words then never straddle two ranks and
the gathered mask is the plain concatenation of the per-rank words.  No collective is needed on the data path; only
the packed result words are gathered.  Row-shaped results (time of impact, goal-sample counts, the roadmap's
per-source distance tables — sources are independent too, and a shard of 32 is one source block of the kernel) are
sharded the same way and gathered with `gather_rows`.
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world_size: int, align: int = 32) -> tuple[int, int]:
    """[lo, hi) of `rank`'s share of n queries; lo is a multiple of `align`; shares differ by one block of `align`
    at most, and the last non-empty share may additionally end inside a block (so sizes differ by less than 2 x align)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    blocks = (n + align - 1) // align
    per, extra = divmod(blocks, world_size)
    lo_b = rank * per + min(rank, extra)
    hi_b = lo_b + per + (1 if rank < extra else 0)
    return min(lo_b * align, n), min(hi_b * align, n)


def words_for(n: int) -> int:
    return (n + 31) // 32


def predicted_motion_costs(a, b, max_step: float = 0.1) -> np.ndarray:
    """State checks a collision-free motion costs: ceil(equal_weights_distance / max_step) + 1, known before launch
    (src/planning/collision_detection.cpp:88-92, distance.cpp:17-25).  Only used to balance shards, so numpy's libm is
    good enough here — the step counts that decide results are computed on the device / re-derived on the host."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    d = np.linalg.norm(a[:, :3] - b[:, :3], axis=1)
    c = np.clip((a[:, 3:7] * b[:, 3:7]).sum(1), -1.0, 1.0)
    d = d + np.arccos(np.clip(2.0 * c * c - 1.0, -1.0, 1.0)) + np.abs(a[:, 7:] - b[:, 7:]).sum(1)
    return np.ceil(np.nan_to_num(d, nan=0.0, posinf=0.0) / max_step).astype(np.int64) + 1


def balanced_bounds(costs, world_size: int, align: int = 32) -> list[tuple[int, int]]:
    """Contiguous shards [lo, hi) per rank with (nearly) equal total cost; every lo is a multiple of `align`, so result
    words still never straddle two ranks.  costs: one non-negative number per query (e.g. predicted_motion_costs)."""
    costs = np.asarray(costs, np.float64)
    n = len(costs)
    if world_size <= 0:
        raise ValueError("bad world_size")
    blocks = (n + align - 1) // align
    pad = np.zeros(blocks * align)
    pad[:n] = costs
    cum = np.concatenate([[0.0], np.cumsum(pad.reshape(blocks, align).sum(1))])  # cost before block k
    total = cum[-1]
    cuts = [0]
    for r in range(1, world_size):
        k = int(np.searchsorted(cum, total * r / world_size, side="left")) if total > 0 else (blocks * r) // world_size
        cuts.append(min(max(k, cuts[-1]), blocks))
    cuts.append(blocks)
    return [(min(cuts[r] * align, n), min(cuts[r + 1] * align, n)) for r in range(world_size)]


def gather_ragged(local, bounds, rank: int, group=None, per_query_words: bool = True):
    """All-gather results of shards with arbitrary (aligned) bounds, e.g. from balanced_bounds.  per_query_words: `local`
    holds packed result words of this rank's shard (one bit per query); otherwise one row per query."""
    import torch
    import torch.distributed as dist

    sizes = [words_for(hi - lo) if per_query_words else hi - lo for lo, hi in bounds]
    pad = max(sizes) if sizes else 0
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: sizes[rank]] = local[: sizes[rank]]
    out = [torch.empty_like(buf) for _ in bounds]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)])


def gather_words(local_words, n: int, rank: int, world_size: int, group=None):
    """All-gather the packed result words of every rank's shard into the full bitmask (torch.distributed; NCCL on
    GPUs, gloo in the CPU tests).  `local_words` is a uint32/int32 torch tensor of this rank's shard."""
    import torch
    import torch.distributed as dist

    sizes = [words_for(hi - lo) for lo, hi in (shard_range(n, r, world_size) for r in range(world_size))]
    pad = max(sizes) if sizes else 0
    buf = torch.zeros(pad, dtype=torch.int32, device=local_words.device)
    buf[: local_words.numel()] = local_words.view(torch.int32)
    out = [torch.empty_like(buf) for _ in range(world_size)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)])


def gather_rows(local_rows, n: int, rank: int, world_size: int, group=None, align: int = 32):
    """All-gather row-sharded results (one row per query: roadmap distance tables sharded by source, goal-sample counts
    sharded by fruit, toi values ...).  `local_rows`: this rank's [hi - lo, ...] torch tensor for shard_range(n, rank)."""
    import torch
    import torch.distributed as dist

    sizes = [hi - lo for lo, hi in (shard_range(n, r, world_size, align) for r in range(world_size))]
    pad = max(sizes) if sizes else 0
    buf = torch.zeros((pad,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)
    buf[: local_rows.shape[0]] = local_rows
    out = [torch.empty_like(buf) for _ in range(world_size)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)])


def concat_words(per_rank_words, n: int) -> np.ndarray:
    """Host-side equivalent of gather_words for already-collected shards (used by tests)."""
    out = np.concatenate([np.asarray(w, dtype=np.uint32) for w in per_rank_words]) if per_rank_words else np.zeros(0, np.uint32)
    assert len(out) == words_for(n)
    return out
