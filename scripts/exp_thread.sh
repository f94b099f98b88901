#!/bin/bash
# usage (GPU box): scripts/exp_thread.sh <lib...>  -- duration of the first 4 dtl_wave_thread_kernel launches per library
for lib in "$@"; do
  echo "== $lib"
  SR2_LIBRARY=$PWD/$lib ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dtl_wave_thread -c 4 --csv \
      python bench.py --profile --steps 1 2>/dev/null | grep -o 'dtl_wave_thread_kernel<[0-9]*>.*' | awk -F'"' '{print $1, $(NF-1)}'
done
