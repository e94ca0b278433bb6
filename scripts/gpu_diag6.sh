#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
rm -f gpurun_out/diag_shear.log
run() { timeout 120 python scripts/diag_clocks.py --n 30 --L 3 "$@" 2>&1 | grep "^n=" | tail -1 >> gpurun_out/diag_shear.log; }
run --opt staged=0 --opt cluster=0 --opt prefetch=1 --opt page_bits=0
run --opt staged=0 --opt cluster=0 --opt prefetch=17 --opt page_bits=0
run --opt staged=0 --opt cluster=1 --opt prefetch=17 --opt page_bits=0
run --opt staged=0 --opt cluster=0 --opt prefetch=0 --opt page_bits=0
run --opt staged=1 --opt cluster=0 --opt page_bits=0
run --opt staged=1 --opt cluster=1 --opt page_bits=0
run --opt staged=3 --opt cluster=0 --opt page_bits=0
cat gpurun_out/diag_shear.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_n30_shear_p1.csv \
    python scripts/prof_run.py --n 30 --L 3 --opt cluster=0 --opt staged=0 --opt prefetch=1 --opt page_bits=0 > gpurun_out/ncu_list.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_n30_shear_st.csv \
    python scripts/prof_run.py --n 30 --L 3 --opt cluster=0 --opt staged=1 --opt page_bits=0 > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_tile12ILi2 -s 0 -c 3 -o gpurun_out/prof_bwd_shear \
    python scripts/prof_run.py --n 28 --L 3 --opt prefetch=1 --opt page_bits=0 --opt cluster=0 > gpurun_out/ncu_full.log 2>&1
