"""Instance sharding across GPUs / ranks.

AECM instances share nothing, so a batch of N instances is split into contiguous ranges, one per rank, with
no data-path collective: rank r of W owns instances [first, first + count).
"""
from __future__ import annotations


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous ceil-split: the first (n_total % world) ranks get one extra instance."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(int(n_total), int(world))
    first = rank * base + min(rank, extra)
    count = base + (1 if rank < extra else 0)
    return first, count


def stream_checksum(out) -> str:
    """SHA-256 of an int16 array's bytes (row-major); used to compare outputs across ranks and runs."""
    import hashlib

    import numpy as np

    return hashlib.sha256(np.ascontiguousarray(out, dtype=np.int16).tobytes()).hexdigest()
