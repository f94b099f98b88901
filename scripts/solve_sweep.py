"""Resident-data solve time of BASELINE config #2 under several settings of the library's switches.
usage (GPU box): PYTHONPATH=. python scripts/solve_sweep.py "VAR=a,b,c" ["VAR2=x;VAR3=y" ...]
Each argument is one sweep: VAR=v1,v2 runs one upload + 8 timed solves per value; NAME=v;NAME2=w sets several at once.
The first line is the default build."""
import os, sys, time, numpy as np, torch
from superrec2_b200 import _lib, synth

batch = synth.make_batch(200, 500, 10000, seed=2000)


def measure(env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        ctx = _lib.Context(0)
        ctx.set_species(batch.parent, batch.child0, batch.child1)
        ctx.dtl_upload(batch.node_off, batch.left, batch.right, batch.leaf_species, batch.costs)
        for _ in range(3):
            ctx.dtl_solve()
        ctx.synchronize()
        times = []
        for _ in range(20):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            ctx.dtl_solve()
            ctx.synchronize()
            b.record()
            torch.cuda.synchronize()
            times.append(ctx.timing().solve_ms if hasattr(ctx.timing(), "solve_ms") else a.elapsed_time(b))
        res = ctx.dtl_results()
        chk = int(np.asarray(res[0]).astype(np.int64).sum())
        ctx.close()
        return float(np.median(times)), float(np.min(times)), chk
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


print("default", measure({}), flush=True)
for arg in sys.argv[1:]:
    if ";" in arg:
        env = dict(kv.split("=", 1) for kv in arg.split(";"))
        print(arg, measure(env), flush=True)
    else:
        name, values = arg.split("=", 1)
        for v in values.split(","):
            print(name, v, measure({name: v}), flush=True)
