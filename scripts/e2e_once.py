"""One pinned-input sr2_dtl_run of BASELINE config #2 (for ncu launch lists of the e2e path)."""
import numpy as np, torch
from superrec2_b200 import _lib, synth
batch = synth.make_batch(200, 500, 10000, seed=2000)
ctx = _lib.Context(0)
ctx.set_species(batch.parent, batch.child0, batch.child1)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
args = (batch.node_off, pin(batch.left), pin(batch.right), pin(batch.leaf_species), batch.costs)
out = tuple(pin(np.zeros(n, np.int32)) for n in (batch.n_instances, batch.n_instances, len(batch.left)))
ctx.dtl_run(*args, out=out)
print("min cost checksum", int(out[0].astype(np.int64).sum()))
