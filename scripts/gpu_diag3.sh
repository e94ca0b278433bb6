#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/diag_page.log
for pb in 17 0 16 18 19 15; do
timeout 300 python scripts/diag_clocks.py --n 30 --L 3 --opt page_bits=$pb 2>&1 | grep "^n=" | tail -1 >> gpurun_out/diag_page.log
done
timeout 300 python scripts/diag_clocks.py --n 28 --L 3 --opt page_bits=17 2>&1 | grep "^n=" | tail -1 >> gpurun_out/diag_page.log
timeout 300 python scripts/diag_clocks.py --n 31 --L 2 --opt page_bits=17 2>&1 | grep "^n=" | tail -1 >> gpurun_out/diag_page.log
timeout 300 python scripts/diag_clocks.py --n 31 --L 2 --opt page_bits=0 2>&1 | grep "^n=" | tail -1 >> gpurun_out/diag_page.log
cat gpurun_out/diag_page.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_n30_lean_pb17.csv \
    python scripts/prof_run.py --n 30 --L 3 --prefetch 1 --lean 3 > gpurun_out/ncu_list.log 2>&1
