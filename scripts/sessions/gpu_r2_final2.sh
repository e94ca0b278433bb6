#!/bin/bash
# round 2: final-tree GPU tier, smoke, one-GPU bench as the driver runs it
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final2_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final2_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final2_smoke.log 2>&1
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final2_bench_n1.log 2>&1
timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/final2_bench_ref.log 2>&1
tail -3 gpurun_out/final2_gpu_tests.log; tail -1 gpurun_out/final2_smoke.log; tail -1 gpurun_out/final2_bench_n1.log | cut -c1-600; tail -1 gpurun_out/final2_bench_ref.log | cut -c1-800
