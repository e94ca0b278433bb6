#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_n20_dram.csv \
    python scripts/prof_run.py --n 20 --L 20 > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_tile12ILi2 -s 3 -c 3 -o gpurun_out/prof_bwd_n26_k11 \
    python scripts/prof_run.py --n 26 --L 2 > gpurun_out/ncu_full.log 2>&1
