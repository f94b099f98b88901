#!/bin/bash
# usage: scripts/prof.sh <label> <kernel-regex> [skip]   (run on the GPU box through gpurun)
# 1. plain run, 2. launch list, 3. one --set full capture of the kernel matching the regex.
set -u
label=$1; regex=$2; skip=${3:-0}
cmd="python bench.py --profile --steps 1"
$cmd > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_${label}.csv $cmd > gpurun_out/ncu.log 2>&1
$cmd > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${regex} -s ${skip} -c 1 \
    -f -o gpurun_out/prof_${label} $cmd > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
