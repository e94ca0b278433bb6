#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/axis10_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/axis10_gpu_tests.log
for sh in 2 8; do for m in 0 15; do
  echo "=== shards=$sh axis_plan=$m" >> gpurun_out/axis10_shards.log
  timeout 600 python scripts/prof_run.py --n 30 --L 6 --reps 2 --seed 1234 --shards $sh --opt axis_plan=$m >> gpurun_out/axis10_shards.log 2>&1
done; done
tail -3 gpurun_out/axis10_gpu_tests.log; cat gpurun_out/axis10_shards.log
