#!/bin/bash
# usage (on the GPU box, through gpurun):
#   scripts/ncu.sh list <label> <command...>                 launch list (gpu__time_duration per launch)
#   scripts/ncu.sh full <label> <kernel-regex> <skip> <command...>   one --set full capture with source
# Each pass runs the plain command first (it must exit 0 without ncu).
set -u
mode=$1; label=$2; shift 2
if [ "$mode" = list ]; then
  "$@" > gpurun_out/plain_${label}.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_${label}.log; exit 1; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
      --log-file gpurun_out/launches_${label}.csv "$@" > gpurun_out/ncu_${label}.log 2>&1
  tail -1 gpurun_out/ncu_${label}.log
else
  regex=$1; skip=$2; shift 2
  "$@" > gpurun_out/plain_${label}.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_${label}.log; exit 1; }
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:${regex}" -s ${skip} -c 1 \
      -f -o gpurun_out/prof_${label} "$@" > gpurun_out/ncu_full_${label}.log 2>&1
  tail -2 gpurun_out/ncu_full_${label}.log
fi
