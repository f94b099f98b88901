import time, numpy as np, os
from superrec2_b200 import _lib, synth
batch = synth.make_batch(200, 500, 10000, seed=2000)
ctx = _lib.Context(0)
ctx.set_species(batch.parent, batch.child0, batch.child1)
args = (batch.node_off, batch.left, batch.right, batch.leaf_species, batch.costs)
for rep in range(4):
    t=time.perf_counter(); ctx.dtl_run(*args); dt=time.perf_counter()-t
    tm=ctx.timing(); print(f"run {dt*1e3:.2f} ms upload {tm.upload_ms:.2f} solve {tm.solve_ms:.2f} waves {tm.waves_ms:.2f} results {tm.results_ms:.2f} launches {tm.n_launches}", flush=True)
