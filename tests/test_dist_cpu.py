"""Multi-process host logic of the sharded paths on CPU (gloo, world size 2): shard ranges
cover the work exactly once and the 8-byte key min-reduce picks the lexicographic minimum
(cost, resolution, root) -- the reference's first strict minimum."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from superrec2_b200 import dist as sr2_dist


def test_shard_ranges_partition_the_work():
    for total in (0, 1, 7, 27, 99225):
        for world in (1, 2, 3, 8):
            blocks = [sr2_dist.shard_range(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))


def test_key_packing_is_lexicographic():
    keys = [sr2_dist.pack_key(c, r, k) for c, r, k in [(5, 9, 1), (5, 2, 7), (4, 30, 30), (5, 2, 3)]]
    assert sr2_dist.unpack_key(min(keys)) == (4, 30, 30)
    assert sr2_dist.unpack_key(sorted(keys)[1]) == (5, 2, 3)


def _worker(rank, world, port, keys, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = sr2_dist.allreduce_min_key(keys[rank])
        results[rank] = out
        # ownership rule used by usreconcile_extended_uspfs_sharded
        results[world + rank] = int(out == keys[rank])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("keys", [
    [sr2_dist.pack_key(16, 5, 8), sr2_dist.pack_key(16, 20, 0)],
    [sr2_dist.NO_KEY, sr2_dist.pack_key(3, 1, 2)],
    [sr2_dist.NO_KEY, sr2_dist.NO_KEY],
])
def test_allreduce_min_key_gloo_world2(keys):
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    manager = mp.Manager()
    results = manager.dict()
    mp.spawn(_worker, args=(2, port, keys, results), nprocs=2, join=True)
    assert results[0] == results[1] == min(keys)
    owners = results[2] + results[3]
    assert owners == (2 if keys[0] == keys[1] else 1)
