"""Multi-GPU partitioning of the hot path: the training step is pure data parallel (batch shards + one
gradient all-reduce per step, done by torch DDP over NCCL/NVLink exactly as Lightning's `strategy: ddp` does); sampling
is embarrassingly parallel over tiles -- a static partition with no collective."""
from __future__ import annotations

from typing import Tuple


def per_rank_batch(global_batch: int, world_size: int) -> int:
    """`PairedDataModule.setup` rule (src/data/paired_data_module.py:273-278): `data.batch_size` is the GLOBAL batch."""
    if global_batch % world_size != 0:
        raise RuntimeError(f"Batch size ({global_batch}) is not divisible by the number of devices ({world_size}).")
    return global_batch // world_size


def shard_range(n_tiles: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of `n_tiles` sampling tiles for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_tiles, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
