import time, sys
sys.path.insert(0, '/root/repo')
from superrec2_b200 import _lib, synth
from bench import resolution_workload
data = synth.make_polytomy_input(arities=(6, 6))
srec_input, species, poly, bits, costs = resolution_workload(data)
total = srec_input.count_resolutions()
print("resolutions", total, "leaves", len(data["leaf_object_species"]))
ctx = _lib.default_context(0)
ctx.set_species(species.parent, species.child0, species.child1)
for rep in range(3):
    t = time.perf_counter()
    ctx.resolutions_upload(poly.child_off, poly.child_idx, poly.leaf_species, bits, costs, 0, total)
    t1 = time.perf_counter()
    ctx.superdtl_solve()
    t2 = time.perf_counter()
    print(f"upload {1e3*(t1-t):.1f} ms solve {1e3*(t2-t1):.1f} ms", ctx.timing().solve_ms, ctx.timing().waves_ms, ctx.superdtl_best()[:3])
