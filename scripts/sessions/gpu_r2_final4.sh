#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final4_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final4_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final4_smoke.log 2>&1
tail -3 gpurun_out/final4_gpu_tests.log; tail -1 gpurun_out/final4_smoke.log
