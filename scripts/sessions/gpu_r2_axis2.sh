#!/bin/bash
# round 2: per-pass device times of static / axis-aware plans run back to back, with clocks and power sampled
mkdir -p gpurun_out
nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,power.limit,temperature.gpu,clocks_throttle_reasons.active --format=csv -lms 250 > gpurun_out/axis2_smi.csv 2>&1 &
SMI=$!
for mode in 0 1; do
  echo "=== axis_plan=$mode" >> gpurun_out/axis2_trace.log
  date +%T.%N >> gpurun_out/axis2_trace.log
  QR_TRACE_PASSES=1 timeout 600 python scripts/prof_run.py --n 30 --L 6 --reps 3 --seed 1234 --opt axis_plan=$mode >> gpurun_out/axis2_trace.log 2>&1
  date +%T.%N >> gpurun_out/axis2_trace.log
done
kill $SMI
tail -60 gpurun_out/axis2_trace.log
