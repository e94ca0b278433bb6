#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1 2 3; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_dbg$dbg.csv \
    python scripts/prof_run.py --n 30 --L 3 --opt debug=$dbg > gpurun_out/ncu_list.log 2>&1
done
