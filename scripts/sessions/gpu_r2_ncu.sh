#!/bin/bash
# Round 2 ncu evidence: launch lists (duration + DRAM bytes per launch) and --set full captures of the hot kernels.
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
# one-GPU gradient, 30 qubits x 3 layers
timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_mcclean30_L3.csv python scripts/prof_run.py --n 30 --L 3 > gpurun_out/ncu1.log 2>&1
# the same circuit through the swap engine on 2 and 8 virtual shards (k_tile12_x exchange passes)
timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_swap_30_L3_g1.csv python scripts/prof_run.py --n 30 --L 3 --shards 2 > gpurun_out/ncu2.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -c 800 --csv --log-file gpurun_out/r2_launches_swap_30_L3_g3.csv python scripts/prof_run.py --n 30 --L 3 --shards 8 > gpurun_out/ncu3.log 2>&1
# full captures: the three backward passes of a layer (one GPU), and backward exchange passes of the swap engine
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_tile12ILi2 -s 3 -c 3 -o gpurun_out/r2_prof_bwd_n30 \
    python scripts/prof_run.py --n 30 --L 3 > gpurun_out/ncu4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_tile12_xILi2 -s 0 -c 2 -o gpurun_out/r2_prof_xbwd_g1 \
    python scripts/prof_run.py --n 30 --L 3 --shards 2 > gpurun_out/ncu5.log 2>&1
ls -la gpurun_out | tail -8
tail -2 gpurun_out/ncu1.log gpurun_out/ncu5.log
