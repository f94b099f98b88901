"""Kernel timeline (CUPTI) of BASELINE config #5: 99,225 resolutions through sr2_resolutions_upload + solve."""
import json, time, numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from superrec2_b200 import _lib, synth
from superrec2_b200.compute.flatten import flatten_polytomies, flatten_species
from superrec2_b200.compute.reconciliation import cost_vector
from superrec2_b200.compute.unordered_super_reconciliation import family_index, leaf_bitsets
from superrec2_b200.model.reconciliation import SuperReconciliationInput
data = synth.make_polytomy_input()
srec_input = SuperReconciliationInput.from_dict(data)
ctx = _lib.default_context(0)
species = flatten_species(srec_input.species_lca.tree)
poly = flatten_polytomies(srec_input.object_tree, srec_input.leaf_object_species, species.rank)
families = family_index(srec_input.leaf_syntenies)
bits = leaf_bitsets(poly.nodes, srec_input.leaf_syntenies, families)
ctx.set_species(species.parent, species.child0, species.child1)
costs = cost_vector(srec_input.costs)
def run():
    ctx.resolutions_upload(poly.child_off, poly.child_idx, poly.leaf_species, bits, costs, 0, 99225)
    ctx.superdtl_solve()
    return ctx.superdtl_best()
for _ in range(3):
    run()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    t = time.perf_counter(); run(); torch.cuda.synchronize(); print("wall ms", 1e3 * (time.perf_counter() - t))
prof.export_chrome_trace("gpurun_out/config5_trace.json")
ev = [e for e in json.load(open("gpurun_out/config5_trace.json"))["traceEvents"]
      if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"]); t0 = ev[0]["ts"]
agg = {}
for e in ev:
    n = e["name"].split("(")[0][:44]; a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += e["dur"]
print("device span ms", (max(e["ts"] + e["dur"] for e in ev) - t0) / 1e3)
for n, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{n:46s} {c:4d} {d / 1e3:8.3f} ms")
for e in ev:
    if e["dur"] > 60:
        print(f"{(e['ts'] - t0) / 1e3:8.3f} +{e['dur'] / 1e3:7.3f} {e['name'].split('(')[0][:50]}")
