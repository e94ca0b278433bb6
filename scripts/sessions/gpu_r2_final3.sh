#!/bin/bash
# round 2: launch list (duration + DRAM bytes) of the final source for roofline.traffic, and the final one-GPU bench line
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_final_mcclean30_L3.csv python scripts/prof_run.py --n 30 --L 3 --seed 1234 > gpurun_out/ncu1.log 2>&1
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final3_bench_n1.log 2>&1
tail -1 gpurun_out/final3_bench_n1.log | cut -c1-400
