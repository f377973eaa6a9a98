"""Multi-GPU plumbing (one process per GPU, torch.distributed). The path only shards where the reference's
arithmetic allows it: evaluation rows are independent -> contiguous row blocks per rank, the item
table replicated, and ONE all-reduce of the six metric sums at the end (no data-path collective)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of n rows for `rank`; block sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_sums(sums: torch.Tensor) -> torch.Tensor:
    """Sum a small tensor of metric sums over ranks (NCCL on GPUs, gloo on CPU). No-op without a process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        sums = sums.clone()
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    return sums
