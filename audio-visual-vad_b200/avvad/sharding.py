"""Utterance sharding for multi-GPU inference: no collective, every rank processes a contiguous block of the
utterance list -- the same partition the reference makes with ``np.array_split(list, n_devices)``
(scripts/evaluate_AV_net.py:329-332)."""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_bounds(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[start, stop) of rank's block; the first n_items % world_size ranks get one extra item (array_split)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_list(items: Sequence, world_size: int, rank: int) -> List:
    a, b = shard_bounds(len(items), world_size, rank)
    return list(items[a:b])


def batches(items: Sequence, batch_size: int):
    """Consecutive call groups of one rank's shard.  The MCB L2 normalisation is per forward call
    (packages/models/AV_Net.py:117), so results depend on this grouping exactly as in the reference."""
    for i in range(0, len(items), batch_size):
        yield list(items[i:i + batch_size])
