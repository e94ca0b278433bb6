#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/diag_pf.log
run() { timeout 120 python scripts/diag_clocks.py --n 30 --L 3 "$@" 2>&1 | grep "^n=" | tail -1 >> gpurun_out/diag_pf.log; }
run --opt prefetch=1
run --opt prefetch=2
run --opt prefetch=3
run --opt prefetch=5
run --opt prefetch=9
cat gpurun_out/diag_pf.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_pf2.csv \
    python scripts/prof_run.py --n 30 --L 3 --prefetch 2 > gpurun_out/ncu_list.log 2>&1
