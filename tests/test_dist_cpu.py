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


def test_packed_key_rejects_fields_that_do_not_fit():
    """no silent truncation of the 23/24/16-bit fields."""
    for bad in [(1 << 23, 0, 0), (0, 1 << 24, 0), (0, 0, 1 << 16), (-1, 0, 0)]:
        assert not sr2_dist.fits_packed_key(bad)
        with pytest.raises(OverflowError):
            sr2_dist.pack_key(*bad)
    assert sr2_dist.fits_packed_key(None) and sr2_dist.fits_packed_key(((1 << 23) - 1, (1 << 24) - 1, 65535))


def _worker(rank, world, port, founds, failed, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        winner, any_failed = sr2_dist.allreduce_min_found(founds[rank], failed=failed[rank])
        results[rank] = (winner, any_failed)
        # ownership rule used by usreconcile_extended_uspfs_sharded
        results[world + rank] = int(winner is not None and winner == founds[rank])
    finally:
        dist.destroy_process_group()


def _spawn(founds, failed=(False, False)):
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    manager = mp.Manager()
    results = manager.dict()
    mp.spawn(_worker, args=(2, port, list(founds), list(failed), results), nprocs=2, join=True)
    return results


@pytest.mark.parametrize("founds", [
    [(16, 5, 8), (16, 20, 0)],
    [None, (3, 1, 2)],
    [None, None],
    [(7, 3, 1), (7, 3, 1)],
    # resolution numbers beyond the 24-bit field (products of two large polytomies): the
    # all-gather fallback must still return the lexicographic minimum
    [(9, (1 << 24) + 5, 2), (9, 77, 40)],
    [(9, (1 << 30) + 5, 2), (9, (1 << 30) + 4, 40)],
])
def test_allreduce_min_found_gloo_world2(founds):
    results = _spawn(founds)
    present = [item for item in founds if item is not None]
    want = min(present) if present else None
    assert results[0] == results[1] == (want, False)
    owners = results[2] + results[3]
    assert owners == (0 if want is None else 2 if founds[0] == founds[1] else 1)


def test_a_failed_rank_does_not_hang_the_collective():
    """a rank whose local solve raised still joins the reduce and everyone learns."""
    results = _spawn([(5, 1, 1), None], failed=(False, True))
    assert results[0] == results[1] == (None, True)
