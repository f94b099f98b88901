#!/bin/bash
# usage: scripts/e2e_sweep.sh  -- e2e ms/step of config #2 (16-bit and 32-bit calls) for pipeline chunk counts /
# first-chunk divisors
export PYTHONPATH=$PWD
for c in 2 3 4; do for f in 1 2 4; do
  SR2_PIPELINE_CHUNKS=$c SR2_PIPELINE_FIRST=$f python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('chunks $c first 1/$f: e2e16', round(d['e2e']['ms_per_step'],3), 'e2e32', round(d['e2e']['int32_ms_per_step'],3), 'pageable', round(d['e2e']['pageable_ms_per_step'],3), 'value', round(d['ms_per_step'],3))"
done; done
