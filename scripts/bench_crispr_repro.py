#!/usr/bin/env python3
"""Bisect why the crispr entry point is slow inside bench.py: stages of bench.py's headline part are
switched on one by one (argv), then the entry point is timed call by call."""
import gzip, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from superrec2_b200 import _lib, synth
import bench

stages = set(sys.argv[1:])
batch = synth.make_batch(200, 500, 10000, 2000)
inputs = (batch.node_off, batch.left, batch.right, batch.leaf_species, batch.costs)
ctx = _lib.Context(0)
if "stream" in stages:
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
ctx.set_species(batch.parent, batch.child0, batch.child1)
ctx.dtl_upload(*inputs)
sampler = bench.ClockSampler(0)
if "sampler" in stages:
    with sampler:
        for _ in range(5):
            ctx.dtl_solve()
else:
    for _ in range(5):
        ctx.dtl_solve()
res = ctx.dtl_results()
if "pinned" in stages:
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    pinputs = (inputs[0], pin(inputs[1]), pin(inputs[2]), pin(inputs[3]), inputs[4])
    out = tuple(pin(np.zeros_like(a)) for a in res)
    for _ in range(3):
        ctx.dtl_run(*pinputs, out=out)
if "pageable" in stages:
    for _ in range(3):
        ctx.dtl_run(*inputs)
if "oracle" in stages:
    bench.time_cpu_port(batch, 96)
ctx.close()

import scripts.entry_timeline  # noqa: F401  (times the entry point call by call)
