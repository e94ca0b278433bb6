#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
rm -f gpurun_out/diag_staged.log
run() { timeout 120 python scripts/diag_clocks.py --n 30 --L 3 "$@" 2>&1 | grep "^n=" | tail -1 >> gpurun_out/diag_staged.log; }
run --opt staged=0 --opt cluster=0 --opt prefetch=17 --opt page_bits=0
run --opt staged=1 --opt cluster=0 --opt page_bits=0
run --opt staged=1 --opt cluster=1 --opt page_bits=0
run --opt staged=1 --opt cluster=2 --opt page_bits=0
run --opt staged=3 --opt cluster=1 --opt page_bits=0
run --opt staged=3 --opt cluster=5 --opt page_bits=0
run --opt staged=1 --opt cluster=1 --opt page_bits=17
cat gpurun_out/diag_staged.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_n30_staged_c0.csv \
    python scripts/prof_run.py --n 30 --L 3 --opt cluster=0 --opt staged=3 --opt page_bits=0 > gpurun_out/ncu_list.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_n30_staged_c2.csv \
    python scripts/prof_run.py --n 30 --L 3 --opt cluster=10 --opt staged=3 --opt page_bits=0 > gpurun_out/ncu_list.log 2>&1
