#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/diag_y.log
run() { timeout 120 python scripts/diag_clocks.py --n 30 --L 3 "$@" 2>&1 | grep "^n=" | tail -1 >> gpurun_out/diag_y.log; }
run --opt staged=4
run --opt staged=12
run --opt staged=8
run --opt staged=4
run --opt staged=12
cat gpurun_out/diag_y.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_y.csv \
    python scripts/prof_run.py --n 30 --L 3 --opt staged=12 > gpurun_out/ncu_list.log 2>&1
