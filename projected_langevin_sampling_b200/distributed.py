"""Multi-GPU partitioning of the Langevin step: one process per GPU, torch.distributed for the plumbing.

Particles are independent given (X, y, Z, V~, lambda) -- every term of the update is column-wise
(reference: src/projected_langevin_sampling/basis/orthonormal.py:151-158) -- so the default partition shards the
particle axis J and needs NO per-step communication; the Philox noise is keyed on the global particle index, so the
result does not depend on the number of GPUs.  When N is too large for one GPU the training rows are sharded as well:
ranks form an (n_groups x j_groups) grid, each rank holds a row slice and a particle slice, and the (M x J_local)
gradient k(Z, X_local) Dc is summed over the ranks that share a particle slice (one all-reduce per step, NCCL over
NVLink) before every member applies the identical update.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional, Tuple

import torch


def shard_range(total: int, rank: int, world: int, align: int = 1) -> Tuple[int, int]:
    """Contiguous [begin, end) of `total` items owned by `rank`; sizes differ by at most `align` units, boundaries are
    multiples of `align` (except the end of the last shard)."""
    units = (total + align - 1) // align
    base, extra = divmod(units, world)
    begin_u = rank * base + min(rank, extra)
    end_u = begin_u + base + (1 if rank < extra else 0)
    return min(begin_u * align, total), min(end_u * align, total)


@dataclass
class GridPlacement:
    """Position of a rank in the (row shards x particle shards) grid; ranks with the same `j_index` form a row group."""

    rank: int
    world: int
    n_groups: int
    j_groups: int

    def __post_init__(self):
        if self.n_groups * self.j_groups != self.world:
            raise ValueError(f"grid {self.n_groups} x {self.j_groups} does not match world size {self.world}")

    @property
    def n_index(self) -> int:
        return self.rank // self.j_groups

    @property
    def j_index(self) -> int:
        return self.rank % self.j_groups

    def rows(self, n_total: int) -> Tuple[int, int]:
        return shard_range(n_total, self.n_index, self.n_groups, align=128)

    def particles(self, j_total: int) -> Tuple[int, int]:
        return shard_range(j_total, self.j_index, self.j_groups, align=2)

    def row_group_ranks(self):
        return [n * self.j_groups + self.j_index for n in range(self.n_groups)]


def make_row_group(placement: GridPlacement):
    """torch.distributed process group of the ranks that share this rank's particle slice (None if rows are not sharded).
    Every rank must call this (new_group is collective)."""
    import torch.distributed as dist

    if placement.n_groups == 1:
        return None
    mine = None
    for j in range(placement.j_groups):
        ranks = [n * placement.j_groups + j for n in range(placement.n_groups)]
        grp = dist.new_group(ranks=ranks)
        if j == placement.j_index:
            mine = grp
    return mine


def gradient_allreduce(group) -> Optional[Callable[[torch.Tensor], None]]:
    """The `gradient_reduce` hook of OrthonormalBasis: sum the (M, J_local) gradient over the row group."""
    if group is None:
        return None
    import torch.distributed as dist

    def reduce_(g: torch.Tensor) -> None:
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)

    return reduce_
