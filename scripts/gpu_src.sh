#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do
for so in 3 0; do
timeout 200 python scripts/diag_clocks.py --n 30 --L 3 --opt src_order=$so 2>&1 | grep "^n=" | tail -1
done
done
timeout 200 python scripts/diag_clocks.py --n 26 --L 6 --opt src_order=3 2>&1 | grep "^n=" | tail -1
timeout 200 python scripts/diag_clocks.py --n 26 --L 6 --opt src_order=0 2>&1 | grep "^n=" | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_src.csv \
    python scripts/prof_run.py --n 30 --L 3 > gpurun_out/ncu_list.log 2>&1
