#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_pair.csv \
    python scripts/prof_run.py --n 30 --L 3 --tile-bits 12 --opt pair=3 > gpurun_out/ncu_list.log 2>&1
