#!/bin/bash
mkdir -p gpurun_out
for m in 13 15; do
  echo "=== axis_plan=$m" >> gpurun_out/axis6_trace.log
  QR_TRACE_PASSES=1 timeout 600 python scripts/prof_run.py --n 30 --L 6 --reps 2 --seed 1234 --opt axis_plan=$m 2>&1 | tail -37 >> gpurun_out/axis6_trace.log
done
timeout 900 python scripts/ab_axis_plan.py --cases 30x30,28x30,26x30,24x30,20x20 --modes 0,13,15 --tile-bits 0,12 --reps 2 > gpurun_out/axis6_ab.log 2>&1
cat gpurun_out/axis6_ab.log
