"""Device timeline of one pinned-input sr2_dtl_run of BASELINE config #2 (CUPTI via torch.profiler):
per stream, the busy time and the first/last activity, then the 40 largest idle gaps of the whole device."""
import json, sys, numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from superrec2_b200 import _lib, synth
batch = synth.make_batch(200, 500, 10000, seed=2000)
ctx = _lib.Context(0)
ctx.set_species(batch.parent, batch.child0, batch.child1)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
args = (batch.node_off, pin(batch.left), pin(batch.right), pin(batch.leaf_species), batch.costs)
out = tuple(pin(np.zeros(n, np.int32)) for n in (batch.n_instances, batch.n_instances, len(batch.left)))
for _ in range(3):
    ctx.dtl_run(*args, out=out)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    ctx.dtl_run(*args, out=out)
    torch.cuda.synchronize()
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/e2e_trace.json"
prof.export_chrome_trace(path)
events = [e for e in json.load(open(path))["traceEvents"]
          if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
events.sort(key=lambda e: e["ts"])
t0 = events[0]["ts"]
end = max(e["ts"] + e["dur"] for e in events)
print(f"device span {(end - t0) / 1e3:.3f} ms, {len(events)} activities")
streams = {}
for e in events:
    streams.setdefault(e["args"].get("stream"), []).append(e)
for s, evs in sorted(streams.items(), key=lambda kv: kv[1][0]["ts"]):
    busy = sum(e["dur"] for e in evs)
    names = {}
    for e in evs:
        n = e["name"].split("(")[0][:40]
        names[n] = names.get(n, 0) + e["dur"]
    top = sorted(names.items(), key=lambda kv: -kv[1])[:4]
    print(f"stream {s}: {len(evs)} acts, busy {busy / 1e3:.3f} ms, first {(evs[0]['ts'] - t0) / 1e3:.3f}, "
          f"last end {(evs[-1]['ts'] + evs[-1]['dur'] - t0) / 1e3:.3f}; "
          + ", ".join(f"{n} {d / 1e3:.2f}" for n, d in top))
# idle gaps on the two compute streams taken together (kernels only)
kern = [e for e in events if e["cat"] == "kernel"]
cur = kern[0]["ts"]
gaps = []
for e in kern:
    if e["ts"] > cur:
        gaps.append((e["ts"] - cur, (cur - t0) / 1e3, e["name"][:40]))
    cur = max(cur, e["ts"] + e["dur"])
print(f"first kernel at {(kern[0]['ts'] - t0) / 1e3:.3f} ms; last kernel ends {(cur - t0) / 1e3:.3f} ms; "
      f"no-kernel gaps total {sum(g[0] for g in gaps) / 1e3:.3f} ms")
for g in sorted(gaps, reverse=True)[:15]:
    print(f"  gap {g[0] / 1e3:.3f} ms at {g[1]:.3f} ms before {g[2]}")
# timeline in 0.5 ms bins: which kernels
for e in events:
    if e["cat"] != "kernel" or e["dur"] > 150:
        print(f"  {(e['ts'] - t0) / 1e3:8.3f} +{e['dur'] / 1e3:6.3f} s{e['args'].get('stream')} {e['name'][:50]} "
              f"{e['args'].get('bytes', '')}")
