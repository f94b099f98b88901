#!/usr/bin/env python3
"""Full-size oracle fixtures for BASELINE.json configs #2-#5 (TEST INFRASTRUCTURE).

    python oracle/make_fullsize_golden.py [2 3 4 5]        (about 10 minutes on 8 cores)

Runs the CPU oracle (``oracle/sr2_oracle.c``, itself pinned to the unmodified reference by
``tests/test_oracle_golden.py``) on the full BASELINE workloads and writes
``tests/golden/fullsize_config<N>.npz``: results of every instance, and one checksum per table row
for the two big single instances (see ``oracle/fullsize.py``).  The ``-m gpu`` tests compare the CUDA
path with these files, so no BASELINE size rests on property checks alone.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import binding as oracle  # noqa: E402
from oracle import fullsize  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
THREADS = os.cpu_count() or 1


def save(name, **arrays):
    path = os.path.join(GOLDEN, name)
    np.savez_compressed(path, **arrays)
    print(f"wrote {path} ({os.path.getsize(path)} bytes)")


def config2():
    batch = fullsize.config2()
    start = time.time()
    min_cost, root, mapping = oracle.dtl_batch(
        batch.parent, batch.child0, batch.child1, batch.node_off, batch.left, batch.right,
        batch.leaf_species, batch.costs, want_map=True, threads=THREADS)
    print(f"config 2: {batch.n_instances} trees in {time.time() - start:.0f} s")
    size = int(batch.node_off[1])
    save("fullsize_config2.npz", min_cost=min_cost, root=root.astype(np.int16),
         map_hash=fullsize.row_checksums(mapping.reshape(-1, size)), checksum=np.int64(min_cost.sum()))


def config3():
    batch = fullsize.config3()
    oracle.load().oracle_set_threads(THREADS)
    start = time.time()
    out = oracle.dtl(batch.parent, batch.child0, batch.child1, batch.left, batch.right, batch.leaf_species,
                     batch.costs, strict=False, tables=True)
    print(f"config 3: {time.time() - start:.0f} s, min cost {out['min_cost']}")
    assert out["rc"] == 0
    save("fullsize_config3.npz", min_cost=np.int32(out["min_cost"]), root=np.int32(out["root_species"]),
         mapping=out["mapping"].astype(np.int16), value_rows=fullsize.row_checksums(out["value"], 1),
         tag_rows=fullsize.row_checksums(out["tag"], 2),
         finite_cells=np.int64((out["value"] < oracle.INF).sum()))


def config4():
    batch = fullsize.config4()
    oracle.load().oracle_set_threads(THREADS)
    start = time.time()
    out = oracle.superdtl(batch.parent, batch.child0, batch.child1, batch.left, batch.right,
                          batch.leaf_species, batch.leaf_family_bits, batch.costs, tables=True)
    print(f"config 4: {time.time() - start:.0f} s, min cost {out['min_cost']}")
    assert out["rc"] == 0
    save("fullsize_config4.npz", min_cost=np.int32(out["min_cost"]), root=np.int32(out["root_species"]),
         mapping=out["mapping"].astype(np.int16), kind=out["kind"].astype(np.int8), syn=out["syn"],
         gain=out["gain"], lca=out["lca"], value_rows=fullsize.row_checksums(out["value"], 1),
         tag_rows=fullsize.row_checksums(out["tag"], 2),
         finite_cells=np.int64((out["value"] < oracle.INF).sum()))


def config5():
    res = fullsize.ResolutionSet(fullsize.config5())
    start = time.time()
    min_cost = np.empty(res.total, np.int32)
    root = np.empty(res.total, np.int32)
    step = 4096
    for first in range(0, res.total, step):
        count = min(step, res.total - first)
        min_cost[first:first + count], root[first:first + count] = oracle.superdtl_batch(
            *res.species_args, *res.batch(first, count), res.costs, threads=THREADS)
    print(f"config 5: {res.total} resolutions in {time.time() - start:.0f} s")
    # the library's unranking against the host model's (reference order) on a sample of resolutions
    from tests.helpers import FlatSuper
    rng = np.random.default_rng(5)
    for index in sorted(rng.choice(res.total, 48, replace=False)):
        binary = res.srec_input.resolution(int(index))
        binary.label_internal()
        flat = FlatSuper(binary)
        out = oracle.superdtl(*flat.species_args, *flat.object_args, flat.leaf_bits, flat.costs, tables=False)
        assert (out["min_cost"], out["root_species"]) == (min_cost[index], root[index]), index
    finite = root >= 0
    order = np.lexsort((root, np.arange(res.total), np.where(finite, min_cost, oracle.INF)))
    best = int(order[0])
    save("fullsize_config5.npz", min_cost=min_cost.astype(np.int16), root=root.astype(np.int16),
         best=np.asarray([min_cost[best], best, root[best]], np.int64))


if __name__ == "__main__":
    which = [int(a) for a in sys.argv[1:]] or [4, 5, 3, 2]
    for number in which:
        {2: config2, 3: config3, 4: config4, 5: config5}[number]()
