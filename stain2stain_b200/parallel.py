"""Multi-GPU partitioning of the hot path: the training step is pure data parallel (batch shards + one
gradient all-reduce per step, done by torch DDP over NCCL/NVLink exactly as Lightning's `strategy: ddp` does); sampling
is embarrassingly parallel over tiles -- a static partition with no collective."""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch


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


class FlatGradients:
    """ONE flat fp32 buffer holding every parameter's gradient back to back (parameter order), plus a view per parameter.

    The data-parallel training step all-reduces this buffer once (SUM) and lets the optimizer read the views with
    `grad_scale = 1 / world` -- the mean `DistributedDataParallel` would have produced with its buckets.
    Device-agnostic host logic (the CPU gloo test exercises it); on CUDA the gather itself is one multi-tensor launch
    (graphed.py, s2s_copy_multi), here `gather()` is the plain torch spelling of the same copy."""

    def __init__(self, params: List[torch.Tensor]):
        self.params = list(params)
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=self.params[0].device)
        self.views: Dict[torch.Tensor, torch.Tensor] = {}
        self.offsets: List[Tuple[int, int]] = []
        off = 0
        for p in self.params:
            self.views[p] = self.flat[off:off + p.numel()].view_as(p)
            self.offsets.append((off, p.numel()))
            off += p.numel()

    def gather(self):
        for p in self.params:
            self.views[p].copy_(p.grad)

    def all_reduce_sum(self, group=None):
        import torch.distributed as dist
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
