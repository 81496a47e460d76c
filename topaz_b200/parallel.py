"""Multi-GPU partitioning of the hot path: one process per GPU (torch.distributed / NCCL for the plumbing).

  * scoring / 2-D denoising: independent micrographs, round-robin over ranks, no data-path collective
    (replaces nothing in the reference: `topaz extract` / `topaz denoise` are single-device);
  * 3-D denoising: independent patches of one tomogram, contiguous ranges per rank, centres pasted on the host
    (replaces torch.nn.DataParallel in topaz/commands/denoise3d.py:103,118, which gives no speed-up at batch 1);
  * GE-binomial training: batch sharding with a logits all-gather + one flat-gradient all-reduce
    (topaz_b200.methods.GE_binomial.step).
"""
from typing import List, Sequence, Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) slice of n work items for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_list(items: Sequence, rank: int, world: int) -> List:
    """Round-robin shard (micrograph i goes to rank i % world); order within a rank is preserved."""
    return list(items[rank::world])
