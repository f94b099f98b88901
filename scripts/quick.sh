#!/bin/bash
# usage: scripts/quick.sh <label>  -- dtl GPU tests + one bench line, short summary
label=$1
python -m pytest tests/test_dtl_gpu.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${label}.json 2>gpurun_out/bench_${label}.err
python -c "
import json;d=json.load(open('gpurun_out/bench_${label}.json'));print('ms/step',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e ms',d['e2e']['ms_per_step'],d['e2e'].get('upload_ms'),d['e2e'].get('results_ms'),'chk',d['checksum_min_costs'])"
tail -3 gpurun_out/bench_${label}.err
