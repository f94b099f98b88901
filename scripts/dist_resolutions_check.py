#!/usr/bin/env python3
"""Multi-GPU check of the sharded-resolutions path (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 scripts/dist_resolutions_check.py

Every rank solves its contiguous block of resolution numbers of BASELINE config #5
(99,225 resolutions), the 8-byte packed key is min-reduced with NCCL, the owning rank
materialises the solution; rank 0 also solves the whole enumeration alone and the two answers
must be identical.  Prints one JSON line with device-event-free wall times (max over ranks)."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from superrec2_b200 import synth  # noqa: E402
from superrec2_b200.compute.unordered_super_reconciliation import usreconcile_extended_uspfs_sharded  # noqa: E402
from superrec2_b200.dist import usreconcile_extended_uspfs_distributed  # noqa: E402
from superrec2_b200.model.reconciliation import SuperReconciliationInput  # noqa: E402
from superrec2_b200.utils.dynamic_programming import RetentionPolicy  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    srec_input = SuperReconciliationInput.from_dict(synth.make_polytomy_input())
    usreconcile_extended_uspfs_distributed(srec_input, RetentionPolicy.ANY, device=local)  # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    start = time.perf_counter()
    got = usreconcile_extended_uspfs_distributed(srec_input, RetentionPolicy.ANY, device=local)
    torch.cuda.synchronize()
    elapsed = torch.tensor([time.perf_counter() - start], device="cuda", dtype=torch.float64)
    owner = torch.tensor([1 if got else 0], device="cuda")
    payload = [json.dumps(next(iter(got)).to_dict()) if got else None]
    if world > 1:
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
        dist.all_reduce(owner, op=dist.ReduceOp.SUM)
        gathered = [None] * world
        dist.all_gather_object(gathered, payload[0])
    else:
        gathered = payload
    if rank == 0:
        assert int(owner.item()) == 1, "exactly one rank must own the winning resolution"
        sharded = next(item for item in gathered if item is not None)
        start = time.perf_counter()
        (alone,) = usreconcile_extended_uspfs_sharded(srec_input, RetentionPolicy.ANY, device=local)
        alone_s = time.perf_counter() - start
        assert json.dumps(alone.to_dict()) == sharded, "sharded answer differs from the single-GPU answer"
        total = srec_input.count_resolutions()
        print(json.dumps({
            "check": "sharded resolutions == single GPU", "world": world, "resolutions": total,
            "min_cost": alone.cost(), "sharded_ms": float(elapsed.item()) * 1e3, "single_ms": alone_s * 1e3,
            "resolutions_per_s": total / float(elapsed.item()),
            "collective": "one all_reduce(MIN) of 8 bytes (NCCL)"}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
