#!/bin/bash
# round 2: GPU tier + estimator timings + the one-GPU bench as the driver runs it
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final_gpu_tests.log
timeout 600 python scripts/time_estimators.py > gpurun_out/final_estimators.log 2>&1
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench_n1.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1
tail -3 gpurun_out/final_gpu_tests.log; cat gpurun_out/final_estimators.log; tail -2 gpurun_out/final_bench_n1.log; tail -1 gpurun_out/final_smoke.log
