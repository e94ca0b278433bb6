#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_n30_lean.csv \
    python scripts/prof_run.py --n 30 --L 3 --prefetch 1 --lean 3 > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
