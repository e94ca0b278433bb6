#!/bin/bash
# round 2, axis-aware plans: GPU tier, one-GPU bench as the driver runs it, launch list with DRAM bytes, full captures of the
# three backward kernels of a layer
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/axis8_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/axis8_gpu_tests.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/axis8_bench_n1.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_axis_mcclean30_L3.csv python scripts/prof_run.py --n 30 --L 3 --seed 1234 > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_tile12 -s 11 -c 3 -o gpurun_out/r2_prof_axis_bwd_n30 \
    python scripts/prof_run.py --n 30 --L 3 --seed 1234 > gpurun_out/ncu4.log 2>&1
tail -3 gpurun_out/axis8_gpu_tests.log; tail -1 gpurun_out/axis8_bench_n1.log | cut -c1-1500
