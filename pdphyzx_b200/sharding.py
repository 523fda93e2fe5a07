"""World-batch sharding across ranks (one process per GPU).

Worlds are independent (reference: no state is shared between PxWorld instances), so a batch
shards by contiguous world-index ranges with no collective on the stepping path.  The only
collectives are a barrier around timed regions, a MAX over per-rank times, and one SUM of a
few statistics at the end of a run.  Works with any torch.distributed backend (NCCL on the
GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_range(n_total: int, rank: int, world_size: int) -> range:
    """Contiguous, balanced partition of [0, n_total): rank r gets ceil/floor shares in order."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, extra = divmod(n_total, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def weak_range(per_rank: int, rank: int) -> range:
    """Weak scaling: every rank owns `per_rank` worlds; global ids stay unique."""
    return range(rank * per_rank, (rank + 1) * per_rank)


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def all_reduce_stats(values: dict, device=None) -> dict:
    """SUM-reduce a dict of scalars over all ranks (identity when not distributed)."""
    dist = _dist()
    if dist is None:
        return dict(values)
    import torch
    keys = sorted(values)
    t = torch.tensor([float(values[k]) for k in keys], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {k: float(v) for k, v in zip(keys, t.cpu().tolist())}


def max_over_ranks(value: float, device=None) -> float:
    dist = _dist()
    if dist is None:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def state_checksum(blocks) -> int:
    """Order-independent 64-bit checksum of a shard's mutable state: the sum over worlds of a
    per-world checksum, so the checksum of a batch equals the sum of its shards' checksums."""
    total = 0
    for w in range(blocks.n):
        nd = int(blocks.header[w]["nDynamic"])
        slots = blocks.order[w, :nd].astype(np.int64)
        words = np.ascontiguousarray(blocks.dyn[w][:, slots]).view(np.uint32).astype(np.uint64)
        weights = (np.arange(words.size, dtype=np.uint64) * np.uint64(2654435761) + np.uint64(1)).reshape(words.shape)
        with np.errstate(over="ignore"):
            total = (total + int((words * weights).sum(dtype=np.uint64))) & 0xFFFFFFFFFFFFFFFF
    return total
