#!/bin/bash
# usage: scripts/e2e_chain_sweep.sh -- e2e ms/step of config #2 for thresholds of the chained top-wave launch
export PYTHONPATH=$PWD
for cr in 0 2368 4736 9472 18944 37888; do
  SR2_DTL_CHAIN_ROWS=$cr python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('chain rows $cr: e2e16', round(d['e2e']['ms_per_step'],3), 'e2e32', round(d['e2e']['int32_ms_per_step'],3), 'value', round(d['ms_per_step'],3))"
done
