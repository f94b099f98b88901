import time, numpy as np
from superrec2_b200 import _lib, synth
from superrec2_b200.compute.flatten import flatten_polytomies, flatten_species
from superrec2_b200.compute.reconciliation import cost_vector
from superrec2_b200.compute.unordered_super_reconciliation import family_index, leaf_bitsets, usreconcile_extended_uspfs
from superrec2_b200.model.reconciliation import SuperReconciliationInput
from superrec2_b200.utils.dynamic_programming import RetentionPolicy
data = synth.make_polytomy_input()
t=time.perf_counter(); srec_input = SuperReconciliationInput.from_dict(data); print("parse", time.perf_counter()-t)
ctx=_lib.default_context(0)
species = flatten_species(srec_input.species_lca.tree)
poly = flatten_polytomies(srec_input.object_tree, srec_input.leaf_object_species, species.rank)
families = family_index(srec_input.leaf_syntenies)
bits = leaf_bitsets(poly.nodes, srec_input.leaf_syntenies, families)
ctx.set_species(species.parent, species.child0, species.child1)
costs=cost_vector(srec_input.costs)
for rep in range(3):
    t=time.perf_counter(); ctx.resolutions_upload(poly.child_off, poly.child_idx, poly.leaf_species, bits, costs, 0, 99225); t1=time.perf_counter()
    ctx.superdtl_solve(); t2=time.perf_counter(); best=ctx.superdtl_best(); t3=time.perf_counter()
    tm=ctx.timing()
    print(f"upload {1e3*(t1-t):.1f} ms  solve {1e3*(t2-t1):.1f} ms (device solve {tm.solve_ms:.1f}, waves {tm.waves_ms:.1f}, launches {tm.n_launches})  best {1e3*(t3-t2):.2f} ms", best[:3])
t=time.perf_counter(); r=usreconcile_extended_uspfs(srec_input, RetentionPolicy.ANY); print("entry total", time.perf_counter()-t)
