"""Partition of the body table over ranks (host logic shared by bench.py and the tests).

Mirrors layout() in csrc/oon_capi.cu: rank g owns the contiguous index block
[g*block, min(n, (g+1)*block)) with block = ceil(n / nranks) rounded up to a whole number of
exchange tiles (TS bodies), so every shard of the all-gathered source buffer starts on a tile
boundary and the source traversal order (hence the result) does not depend on the sharding.
"""
from __future__ import annotations

TS = 256  # bodies per exchange tile == oon::TS in csrc/oon_internal.cuh


CANONICAL_BLOCKS = 8


def block_size(n: int, nranks: int) -> int:
    """Bodies per rank.  The world is cut in 8 canonical blocks (tile-aligned) whatever the GPU count; a rank owns
    8/nranks consecutive ones, so the slot order of the exchange buffer is the same on 1, 2, 4 and 8 GPUs."""
    nblocks = CANONICAL_BLOCKS if nranks in (1, 2, 4, 8) else nranks
    per = -(-n // nblocks)
    cblock = max(TS, -(-per // TS) * TS)
    return cblock * (nblocks // nranks)


def shard_range(n: int, rank: int, nranks: int) -> tuple[int, int]:
    b = block_size(n, nranks)
    lo = min(n, rank * b)
    return lo, min(n, lo + b)


def interactions_per_step(n: int) -> int:
    """directed body-pair interactions of one all-pairs frame"""
    return n * (n - 1)
