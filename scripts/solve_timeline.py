"""Kernel timeline (CUPTI via torch.profiler) of one stepwise sr2_dtl_solve of BASELINE config #2."""
import json, sys, numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from superrec2_b200 import _lib, synth
batch = synth.make_batch(200, 500, 10000, seed=2000)
ctx = _lib.Context(0)
ctx.set_species(batch.parent, batch.child0, batch.child1)
ctx.dtl_upload(batch.node_off, batch.left, batch.right, batch.leaf_species, batch.costs)
for _ in range(3):
    ctx.dtl_solve()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ctx.dtl_solve()
    torch.cuda.synchronize()
path = "gpurun_out/solve_trace.json"
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("ph") == "X" and e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
print("span ms", (max(e["ts"] + e["dur"] for e in ev) - t0) / 1e3)
for e in ev:
    n = e["name"].split("(")[0].replace("dtl_", "").replace("void ", "").replace("_kernel", "")
    print(f"{(e['ts'] - t0) / 1e3:8.3f} +{e['dur']:8.1f}us s{e['args']['stream']} {n} grid {e['args'].get('grid')}")
