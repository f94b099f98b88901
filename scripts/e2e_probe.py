import time, numpy as np, os, torch
from superrec2_b200 import _lib, synth
batch = synth.make_batch(200, 500, 10000, seed=2000)
ctx = _lib.Context(0)
ctx.set_species(batch.parent, batch.child0, batch.child1)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
pageable = (batch.node_off, batch.left, batch.right, batch.leaf_species, batch.costs)
pinned = (batch.node_off, pin(batch.left), pin(batch.right), pin(batch.leaf_species), batch.costs)
out = (np.zeros(batch.n_instances, np.int32), np.zeros(batch.n_instances, np.int32), np.zeros(len(batch.left), np.int32))
for name, args in (("pageable", pageable), ("pinned", pinned)):
    for rep in range(4):
        t=time.perf_counter(); ctx.dtl_run(*args, out=out); dt=time.perf_counter()-t
        tm=ctx.timing(); print(f"{name}: run {dt*1e3:.2f} ms upload {tm.upload_ms:.2f} solve {tm.solve_ms:.2f} waves {tm.waves_ms:.2f} results {tm.results_ms:.2f} launches {tm.n_launches}", flush=True)
