"""Multi-GPU sharding: captures are independent (Receiver.load touches one file,
afskmodem.py:420-430), so a batch is cut into contiguous capture ranges balanced by cumulative
sample count — one range per rank/GPU — with no collective on the data path.  Results (a few
bytes per capture) are gathered on the host at the end.
"""
from __future__ import annotations

import numpy as np


def shard_captures(lengths, world_size: int) -> list[tuple[int, int]]:
    """Contiguous [lo, hi) capture ranges, one per rank, balanced by total samples.

    Rank r gets the captures whose cumulative-sample midpoint falls in the r-th 1/world_size
    slice of the total, so ranges are disjoint, ordered and cover [0, B)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    B = len(lengths)
    if world_size <= 1 or B == 0:
        return [(0, B)] + [(B, B)] * (max(world_size, 1) - 1)
    cum = np.cumsum(lengths)
    total = int(cum[-1])
    mids = cum - lengths / 2.0
    owner = np.minimum((mids * world_size / max(total, 1)).astype(np.int64), world_size - 1)
    owner = np.maximum.accumulate(owner)
    bounds = np.searchsorted(owner, np.arange(world_size + 1), side="left")
    return [(int(bounds[r]), int(bounds[r + 1])) for r in range(world_size)]


def gather_batches(local, group=None):
    """All-gather arbitrary picklable per-rank results to every rank (torch.distributed, any backend).
    Returns the list ordered by rank; with no initialised process group returns [local]."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [local]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, local, group=group)
    return out
