#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
rm -f gpurun_out/diag_piped.log
run() { timeout 120 python scripts/diag_clocks.py "$@" 2>&1 | grep "^n=" | tail -1 >> gpurun_out/diag_piped.log; }
run --n 30 --L 3
run --n 30 --L 3 --opt staged=0
run --n 28 --L 3
run --n 26 --L 6
run --n 20 --L 20
cat gpurun_out/diag_piped.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_piped.csv \
    python scripts/prof_run.py --n 30 --L 3 > gpurun_out/ncu_list.log 2>&1
