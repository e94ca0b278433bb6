#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/diag_lb.log
run() { timeout 120 python scripts/diag_clocks.py --n 30 --L 3 "$@" 2>&1 | grep "^n=" | tail -1 >> gpurun_out/diag_lb.log; }
run --opt low_bits_pass=0
run --opt low_bits_pass=-1
run --opt low_bits_pass=1
run --opt low_bits_pass=0
run --opt low_bits_pass=-1
cat gpurun_out/diag_lb.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_lb.csv \
    python scripts/prof_run.py --n 30 --L 3 --opt low_bits_pass=-1 > gpurun_out/ncu_list.log 2>&1
timeout 200 python scripts/diag_clocks.py --n 26 --L 6 --opt low_bits_pass=0 2>&1 | grep "^n=" | tail -1
timeout 200 python scripts/diag_clocks.py --n 26 --L 6 --opt low_bits_pass=-1 2>&1 | grep "^n=" | tail -1
timeout 200 python scripts/diag_clocks.py --n 20 --L 20 --opt low_bits_pass=0 2>&1 | grep "^n=" | tail -1
timeout 200 python scripts/diag_clocks.py --n 20 --L 20 --opt low_bits_pass=-1 2>&1 | grep "^n=" | tail -1
