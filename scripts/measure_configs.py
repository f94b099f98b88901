#!/usr/bin/env python3
"""Measure the other BASELINE.json configurations (not the bench.py default) on one GPU.

Prints one JSON line per configuration: wall time of the entry-point / C-ABI call and the
library's CUDA-event timings.  Config numbering follows BASELINE.json / SURVEY.md section 8d."""
import json
import sys
import time

import numpy as np

from superrec2_b200 import _lib, synth
from superrec2_b200.compute.unordered_super_reconciliation import usreconcile_extended_uspfs
from superrec2_b200.model.reconciliation import SuperReconciliationInput
from superrec2_b200.utils.dynamic_programming import RetentionPolicy


def timed(fn, repeat=5):
    fn()
    times = []
    for _ in range(repeat):
        start = time.perf_counter()
        out = fn()
        times.append(time.perf_counter() - start)
    return out, min(times) * 1e3, float(np.median(times)) * 1e3


def main():
    ctx = _lib.default_context(0)
    # ---- config #3: dtl, one 5,000-leaf object tree vs a 2,000-leaf species tree -------------
    batch = synth.make_batch(2000, 5000, 1, seed=3000)
    ctx.set_species(batch.parent, batch.child0, batch.child1)
    args = (batch.node_off, batch.left, batch.right, batch.leaf_species, batch.costs)
    out, best, med = timed(lambda: ctx.dtl_run(*args))
    t = ctx.timing()
    print(json.dumps({"config": 3, "what": "dtl 5000-leaf x 2000-leaf, sr2_dtl_run host->host",
                      "cells": batch.cells, "min_cost": int(out[0][0]), "ms_best": best, "ms_median": med,
                      "solve_ms": t.solve_ms, "waves_ms": t.waves_ms, "upload_ms": t.upload_ms,
                      "launches": t.n_launches, "cells_per_s_e2e": batch.cells / (best * 1e-3)}))
    # ---- config #4: superdtl, 1,000-leaf synteny tree, 32 families, 500-leaf species tree ----
    batch = synth.make_batch(500, 1000, 1, seed=4000, n_families=32)
    ctx.set_species(batch.parent, batch.child0, batch.child1)
    args = (batch.node_off, batch.left, batch.right, batch.leaf_species, batch.leaf_family_bits, batch.costs)
    out, best, med = timed(lambda: ctx.superdtl_run(*args))
    t = ctx.timing()
    kinds = batch.cells * 2
    print(json.dumps({"config": 4, "what": "superdtl 1000-leaf x 500-leaf, 32 families, superdtl_run host->host",
                      "cell_kinds": kinds, "min_cost": int(out[0][0]), "ms_best": best, "ms_median": med,
                      "solve_ms": t.solve_ms, "waves_ms": t.waves_ms, "upload_ms": t.upload_ms,
                      "launches": t.n_launches, "cell_kinds_per_s_e2e": kinds / (best * 1e-3)}))
    # ---- config #5: superdtl over the 99,225 resolutions of a 33-leaf multifurcating tree ----
    data = synth.make_polytomy_input()
    srec_input = SuperReconciliationInput.from_dict(data)
    total = srec_input.count_resolutions()
    out, best, med = timed(lambda: usreconcile_extended_uspfs(srec_input, RetentionPolicy.ANY), repeat=3)
    t = ctx.timing()
    (result,) = out
    kinds = total * 32 * 59 * 2
    print(json.dumps({"config": 5, "what": "superdtl entry point, 99,225 resolutions (33 leaves, 59 species, 32 families)",
                      "resolutions": total, "cell_kinds": kinds, "min_cost": result.cost(),
                      "ms_best": best, "ms_median": med, "resolutions_per_s": total / (best * 1e-3),
                      "cell_kinds_per_s_e2e": kinds / (best * 1e-3)}))
    # ---- config #1: crispr is timed by tests; here the example batch of config #2 is in bench.py


if __name__ == "__main__":
    sys.exit(main())
