"""Sharding of a scan sequence across ranks: contiguous blocks of the pair index range.

Pair k registers scan k+1 onto scan k and depends on nothing else, so rank r of W owns pairs
[r*ceil(P/W), (r+1)*ceil(P/W)) of the P = n_scans-1 pairs and needs the scans of those pairs: its own block plus ONE
halo scan (the first scan of the next block).  No feature set ever crosses ranks; there is no data-path collective.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class Shard:
    rank: int
    world: int
    pair_lo: int   # first pair owned
    pair_hi: int   # one past the last pair owned
    scan_lo: int   # first scan needed
    scan_hi: int   # one past the last scan needed (includes the halo scan)

    @property
    def n_pairs(self) -> int:
        return self.pair_hi - self.pair_lo

    @property
    def n_scans(self) -> int:
        return self.scan_hi - self.scan_lo


def shard_sequence(n_scans: int, world: int, rank: int) -> Shard:
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    n_pairs = max(n_scans - 1, 0)
    per = -(-n_pairs // world) if n_pairs else 0
    lo = min(rank * per, n_pairs)
    hi = min(lo + per, n_pairs)
    if hi == lo:  # nothing to do on this rank
        return Shard(rank, world, lo, lo, lo, lo)
    return Shard(rank, world, lo, hi, lo, hi + 1)


def weak_scaling_sequence(scans_per_rank: int, world: int) -> int:
    """Length of the virtual sequence whose shards give every rank `scans_per_rank` pairs (the last rank one fewer)."""
    return scans_per_rank * world
