#!/usr/bin/env python3
"""For ncu: BASELINE config #3 (dtl, one 5,000-leaf object tree vs a 2,000-leaf species tree), one upload + solves."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from superrec2_b200 import _lib, synth  # noqa: E402

solves = int(sys.argv[1]) if len(sys.argv) > 1 else 1
batch = synth.make_batch(2000, 5000, 1, seed=3000)
ctx = _lib.default_context(0)
ctx.set_species(batch.parent, batch.child0, batch.child1)
ctx.dtl_upload(batch.node_off, batch.left, batch.right, batch.leaf_species, batch.costs)
for _ in range(solves):
    ctx.dtl_solve()
t = ctx.timing()
print(f"config 3: solve {t.solve_ms:.3f} ms, waves {t.waves_ms:.3f} ms, {t.n_launches} launches, {t.n_waves} waves")
