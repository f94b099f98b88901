"""Multi-GPU sharding: one process per GPU, ``torch.distributed`` for the plumbing.

The DP shards only along axes without a data-path exchange:

* batches of independent instances -> a contiguous block of instances per rank, no
  collective (``shard_range``);
* enumerated polytomy resolutions -> a contiguous block of resolution numbers per rank,
  then ONE ``all_reduce(MIN)`` of the packed key ``cost<<40 | resolution<<16 | root`` (NCCL over
  NVLink on GPUs, gloo in the CPU tests).  Lexicographic minimum == the reference's first strict
  minimum over (resolution, root species)
  (``/root/reference/src/superrec2/compute/unordered_super_reconciliation.py:454-497``).
"""
from __future__ import annotations

import os
from typing import Optional, Set, Tuple

NO_KEY = (1 << 63) - 1          # "this rank found no solution" (int64 max: loses every MIN)
COST_BITS, RESOLUTION_BITS, ROOT_BITS = 23, 24, 16

Found = Optional[Tuple[int, int, int]]   # (cost, resolution number, root species rank)


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block ``[lo, hi)`` of ``total`` units owned by ``rank``."""
    return total * rank // world, total * (rank + 1) // world


def fits_packed_key(found: Found) -> bool:
    if found is None:
        return True
    cost, resolution, root = found
    return (0 <= cost < 1 << COST_BITS and 0 <= resolution < 1 << RESOLUTION_BITS
            and 0 <= root < 1 << ROOT_BITS)


def pack_key(cost: int, resolution: int, root: int) -> int:
    if not fits_packed_key((cost, resolution, root)):
        raise OverflowError(f"(cost {cost}, resolution {resolution}, root {root}) does not fit the "
                            f"{COST_BITS}/{RESOLUTION_BITS}/{ROOT_BITS}-bit packed key")
    return (cost << 40) | (resolution << 16) | root


def unpack_key(key: int) -> Tuple[int, int, int]:
    return key >> 40, (key >> 16) & 0xFFFFFF, key & 0xFFFF


def allreduce_min_found(found: Found, device=None, failed: bool = False):
    """
    Lexicographic minimum of one ``(cost, resolution, root)`` per rank -> ``(winner, any_failed)``.

    Normal case: ONE all-reduce(MIN) of 16 bytes -- the packed 8-byte key and an "every rank is
    fine" flag, so that a rank whose local solve raised still takes part and nobody is left
    waiting in the collective.  Keys that do not fit the packed fields (resolution numbers of
    products of large polytomies, >= 2^24) take a second step: every rank contributes its three
    numbers to an all-gather and the minimum is taken on the host.
    """
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return found, failed
    packed = fits_packed_key(found)
    key = NO_KEY if found is None or not packed else pack_key(*found)
    # word 0: packed key; word 1: 1 unless this rank failed; word 2: 1 unless a key did not fit
    tensor = torch.tensor([key, 0 if failed else 1, 1 if packed else 0], dtype=torch.int64, device=device)
    dist.all_reduce(tensor, op=dist.ReduceOp.MIN)
    key, all_fine, all_packed = (int(v) for v in tensor.tolist())
    if not all_fine:
        return None, True
    if all_packed:
        return (None if key == NO_KEY else unpack_key(key)), False
    mine = torch.tensor(list(found) if found is not None else [-1, -1, -1], dtype=torch.int64, device=device)
    everyone = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(everyone, mine)
    tuples = [tuple(int(v) for v in item.tolist()) for item in everyone]
    tuples = [item for item in tuples if item[0] >= 0]
    return (min(tuples) if tuples else None), False


def usreconcile_extended_uspfs_distributed(srec_input, policy=None, device: Optional[int] = None,
                                           timings: Optional[dict] = None) -> Set:
    """
    ``superdtl`` with the resolutions of a multifurcating object tree sharded over the ranks
    of the default process group.  Every rank must call it with the same input; the rank that
    owns the optimal resolution returns the solution, the others an empty set.  ``device``
    defaults to ``LOCAL_RANK`` (one process per GPU).  ``timings`` (optional dict) receives the
    wall-clock milliseconds of the three phases and the winning ``(cost, resolution, root)``.
    """
    import time

    import torch
    import torch.distributed as dist

    from . import _lib
    from .compute.unordered_super_reconciliation import materialise_resolution, solve_resolution_range
    from .utils.dynamic_programming import RetentionPolicy

    if policy not in (None, RetentionPolicy.ANY):
        raise NotImplementedError("the sharded path implements RetentionPolicy.ANY")
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", torch.cuda.current_device() if torch.cuda.is_available() else 0))
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    tensor_device = torch.device("cuda", device) if torch.cuda.is_available() else None
    t0 = time.perf_counter()
    error, found = None, None
    try:
        lo, hi = shard_range(srec_input.count_resolutions(), rank, world)
        found = solve_resolution_range(srec_input, lo, hi, device)
    except Exception as exc:  # noqa: BLE001 -- take part in the collective first, then re-raise
        error = exc
    t1 = time.perf_counter()
    winner, any_failed = allreduce_min_found(found, tensor_device, failed=error is not None)
    t2 = time.perf_counter()
    if error is not None:
        raise error
    if any_failed:
        raise _lib.Sr2Error(-5, "another rank failed while solving its block of resolutions")
    result = set()
    if winner is not None and winner == found:
        result = {materialise_resolution(srec_input, winner, device)}
    if timings is not None:
        timings.update(solve_ms=(t1 - t0) * 1e3, reduce_ms=(t2 - t1) * 1e3,
                       materialise_ms=(time.perf_counter() - t2) * 1e3, winner=winner)
    return result
