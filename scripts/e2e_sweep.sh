#!/bin/bash
# usage: scripts/e2e_sweep.sh  -- e2e ms/step of config #2 for pipeline chunk counts / first-chunk divisors
export PYTHONPATH=$PWD
for c in 2 3 4 6; do for f in 1 2 4; do
  SR2_PIPELINE_CHUNKS=$c SR2_PIPELINE_FIRST=$f python bench.py --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('chunks $c first 1/$f: e2e', round(d['e2e']['ms_per_step'],3), 'value', round(d['ms_per_step'],3))"
done; done
