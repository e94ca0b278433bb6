#!/bin/bash
mkdir -p gpurun_out
QR_TRACE_PASSES=1 timeout 600 python scripts/prof_run.py --n 30 --L 6 --reps 2 --seed 1234 2>&1 | tail -37 > gpurun_out/axis12_trace.log
timeout 900 python scripts/ab_axis_plan.py --cases 30x30,28x30,24x30,20x20,16x16 --modes 0,15 --tile-bits 0 --reps 3 > gpurun_out/axis12_ab.log 2>&1
timeout 300 python -m pytest tests/test_parity.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -2 >> gpurun_out/axis12_ab.log
grep "pass 0" gpurun_out/axis12_trace.log | awk '{print $4,$6,$8,$9,$11,$13,$14}' | paste - - - ; cat gpurun_out/axis12_ab.log
