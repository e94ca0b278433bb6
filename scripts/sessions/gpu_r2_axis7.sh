#!/bin/bash
mkdir -p gpurun_out
for opts in "tile_bits=11 axis_plan=0" "tile_bits=11 axis_plan=15"; do
  echo "=== $opts" >> gpurun_out/axis7_trace.log
  o=""; for x in $opts; do o="$o --opt $x"; done
  QR_TRACE_PASSES=1 timeout 600 python scripts/prof_run.py --n 30 --L 6 --reps 2 --seed 1234 $o 2>&1 | tail -61 >> gpurun_out/axis7_trace.log
done
grep -c trace gpurun_out/axis7_trace.log
