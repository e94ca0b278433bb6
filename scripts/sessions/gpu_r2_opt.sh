#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_optimization.py -m gpu -x -q > gpurun_out/opt_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/opt_tests.log
timeout 900 python scripts/opt_loop_bench.py > gpurun_out/opt_loop_bench.log 2>&1
tail -3 gpurun_out/opt_tests.log; cat gpurun_out/opt_loop_bench.log
