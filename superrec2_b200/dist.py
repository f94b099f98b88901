"""Multi-GPU sharding: one process per GPU, ``torch.distributed`` for the plumbing.

The DP shards only along axes without a data-path exchange:

* batches of independent instances -> a contiguous block of instances per rank, no
  collective (``shard_range``);
* enumerated polytomy resolutions -> a contiguous block of resolution numbers per rank,
  then ONE 8-byte ``all_reduce(MIN)`` of the packed key ``cost<<40 | resolution<<16 |
  root`` (NCCL over NVLink on GPUs, gloo in the CPU tests).  Lexicographic minimum ==
  the reference's first strict minimum over (resolution, root species)
  (``/root/reference/src/superrec2/compute/unordered_super_reconciliation.py:454-497``).
"""
from __future__ import annotations

from typing import Set, Tuple

NO_KEY = (1 << 64) - 1


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block ``[lo, hi)`` of ``total`` units owned by ``rank``."""
    return total * rank // world, total * (rank + 1) // world


def pack_key(cost: int, resolution: int, root: int) -> int:
    return (cost << 40) | (resolution << 16) | root


def unpack_key(key: int) -> Tuple[int, int, int]:
    return key >> 40, (key >> 16) & 0xFFFFFF, key & 0xFFFF


def allreduce_min_key(key: int, device=None) -> int:
    """Global minimum of one packed key per rank (8 bytes on the wire)."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return key
    # keys stay below 2^63 (costs < 2^23) except NO_KEY, which maps to int64 max
    value = min(key, (1 << 63) - 1)
    tensor = torch.tensor([value], dtype=torch.int64, device=device)
    dist.all_reduce(tensor, op=dist.ReduceOp.MIN)
    out = int(tensor.item())
    return NO_KEY if out == (1 << 63) - 1 else out


def usreconcile_extended_uspfs_distributed(srec_input, policy=None, device: int = 0) -> Set:
    """
    ``superdtl`` with the resolutions of a multifurcating object tree sharded over the ranks
    of the default process group.  Every rank must call it with the same input; the rank that
    owns the optimal resolution returns the solution, the others an empty set.
    """
    import torch
    import torch.distributed as dist

    from .compute.unordered_super_reconciliation import usreconcile_extended_uspfs_sharded
    from .utils.dynamic_programming import RetentionPolicy

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    tensor_device = torch.device("cuda", device) if torch.cuda.is_available() else None
    return usreconcile_extended_uspfs_sharded(
        srec_input, policy or RetentionPolicy.ANY, device=device, shard=rank, n_shards=world,
        reduce_key=lambda key: allreduce_min_key(key, tensor_device))
