#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for pf in 1 17 2; do
  echo "=== prefetch=$pf" >> gpurun_out/axis9_trace.log
  QR_TRACE_PASSES=1 timeout 600 python scripts/prof_run.py --n 30 --L 6 --reps 2 --seed 1234 --opt prefetch=$pf 2>&1 | tail -37 >> gpurun_out/axis9_trace.log
done
timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/axis9_launches_pf17.csv python scripts/prof_run.py --n 30 --L 2 --seed 1234 --opt prefetch=17 > gpurun_out/ncu1.log 2>&1
grep "nv 2" gpurun_out/axis9_trace.log | awk '{print $4,$6,$9,$11,$13,$14}' | paste - - - - - - | tail -12
