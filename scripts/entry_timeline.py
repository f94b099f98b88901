#!/usr/bin/env python3
"""Where the wall time of the superdtl Python entry point goes: every libsr2 call of
usreconcile_extended_uspfs is timed (host wall clock) for BASELINE configs #1 and #5."""
import gzip
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from superrec2_b200 import _lib, synth  # noqa: E402
from superrec2_b200.compute.unordered_super_reconciliation import usreconcile_extended_uspfs  # noqa: E402
from superrec2_b200.model.reconciliation import SuperReconciliationInput  # noqa: E402
from superrec2_b200.utils.dynamic_programming import RetentionPolicy  # noqa: E402

log = []


def wrap(name):
    inner = getattr(_lib.Context, name)

    def timed(self, *args, **kwargs):
        start = time.perf_counter()
        out = inner(self, *args, **kwargs)
        log.append((name, (time.perf_counter() - start) * 1e3))
        return out
    setattr(_lib.Context, name, timed)


for method in ("set_species", "superdtl_upload", "resolutions_upload", "superdtl_solve", "superdtl_best",
               "superdtl_results", "superdtl_set_allowed"):
    wrap(method)

if "--after-big-batch" in sys.argv:
    # the state bench.py is in when it reaches the secondary configurations: a 10,000-tree dtl batch was
    # solved on another context (17 GB of pools, pinned staging) and that context was closed
    import numpy as np
    import torch
    batch = synth.make_batch(200, 500, 10000, 2000)
    big = _lib.Context(0)
    big.set_species(batch.parent, batch.child0, batch.child1)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    keep = [pin(batch.left), pin(batch.right), pin(batch.leaf_species)]
    big.dtl_run(batch.node_off, *keep, batch.costs)
    if "--keep-open" not in sys.argv:
        big.close()
    for _ in range(3):
        start = time.perf_counter()
        free = torch.cuda.mem_get_info()
        print(f"cudaMemGetInfo {1e3 * (time.perf_counter() - start):.2f} ms, free {free[0] >> 20} MiB", flush=True)

if "--after-oracle" in sys.argv:
    # bench.py runs the CPU oracle (OpenMP, all cores) before the secondary configurations
    from oracle import binding as oracle
    small = synth.make_batch(200, 500, 32, 2000)
    oracle.dtl_batch(small.parent, small.child0, small.child1, small.node_off, small.left, small.right,
                     small.leaf_species, small.costs, want_map=True, threads=os.cpu_count())

with gzip.open(os.path.join(ROOT, "tests", "golden", "crispr.json.gz"), "rt") as handle:
    crispr = json.load(handle)[0]["input"]
for name, data in (("crispr", crispr), ("config5", synth.make_polytomy_input())):
    srec_input = SuperReconciliationInput.from_dict(data)
    for rep in range(4):
        log.clear()
        start = time.perf_counter()
        usreconcile_extended_uspfs(srec_input, RetentionPolicy.ANY)
        total = (time.perf_counter() - start) * 1e3
        inside = sum(ms for _, ms in log)
        print(f"{name} run {rep}: total {total:.2f} ms, in libsr2 {inside:.2f} ms, python {total - inside:.2f} ms :: " +
              ", ".join(f"{n} {ms:.2f}" for n, ms in log), flush=True)
