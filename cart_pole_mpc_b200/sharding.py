"""Multi-GPU sharding of a batch of independent MPC instances.

Instances share nothing, so the data path has NO collective: rank r owns the contiguous range
[r*B/G, (r+1)*B/G) of instance ids and regenerates / loads only its shard.  The only communication is the optional
epilogue all-gather of results (controls and the fixed-size stat records) over NCCL (gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [first, first+count) of `total` instances for `rank`; the first `total % world` ranks get one extra."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(total, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def all_gather_results(local, total: int, group=None):
    """All-gathers a per-instance result tensor [count, ...] into [total, ...] in instance order.

    Works with any torch.distributed backend (nccl on GPUs, gloo in the CPU tests); shards may be ragged."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    counts = [shard_range(total, r, world)[1] for r in range(world)]
    max_count = max(counts)
    tail = tuple(local.shape[1:])
    padded = torch.zeros((max_count,) + tail, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * max_count,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    pieces = [out[r * max_count: r * max_count + counts[r]] for r in range(world)]
    return torch.cat(pieces, dim=0)


def stats_as_bytes(stats: np.ndarray) -> np.ndarray:
    """View a structured cpmpc_stats array as [B, 56] bytes (for all-gather)."""
    return np.frombuffer(stats.tobytes(), dtype=np.uint8).reshape(stats.shape[0], -1)
