"""(batch x heads) data-parallel sharding of the attention forward across GPUs.

Every (batch, head) pair is an independent problem -- the reference already maps them to
independent CTAs (/root/reference/src/include/forward_kernel.cuh:104-118) -- so multi-GPU needs
no collective: each rank owns a rectangle of the (batch, head) index space, holds/generates
only that slice and runs the single-GPU operator on it.  Batch is split first (contiguous in the
(B, N, H, d) layout); heads are split only when world_size does not divide batch.
"""
from __future__ import annotations

from dataclasses import dataclass
from math import gcd


@dataclass(frozen=True)
class Shard:
    rank: int
    world_size: int
    b0: int
    b1: int
    h0: int
    h1: int

    @property
    def batch(self) -> int:
        return self.b1 - self.b0

    @property
    def heads(self) -> int:
        return self.h1 - self.h0

    @property
    def units(self) -> int:
        """(batch, head) problems owned by this rank."""
        return self.batch * self.heads

    def take(self, t):
        """Slice a full (B, N, H, d) tensor down to this shard (a view)."""
        return t[self.b0:self.b1, :, self.h0:self.h1]


def plan_shards(batch: int, n_heads: int, world_size: int):
    """Partition the (batch, head) grid into `world_size` disjoint rectangles of equal size.

    world_size = gb * gh with gb = gcd(world_size, batch) ranks along batch and gh along heads;
    raises ValueError if n_heads is not divisible by gh (no silent imbalance).
    """
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    gb = gcd(world_size, batch)
    gh = world_size // gb
    if n_heads % gh != 0:
        raise ValueError(
            f"cannot split batch={batch} x heads={n_heads} evenly over {world_size} ranks "
            f"(need heads % {gh} == 0)")
    per_b, per_h = batch // gb, n_heads // gh
    shards = []
    for r in range(world_size):
        ib, ih = divmod(r, gh)
        shards.append(Shard(r, world_size, ib * per_b, (ib + 1) * per_b, ih * per_h, (ih + 1) * per_h))
    return shards


def shard_for_rank(batch: int, n_heads: int, world_size: int, rank: int) -> Shard:
    return plan_shards(batch, n_heads, world_size)[rank]
