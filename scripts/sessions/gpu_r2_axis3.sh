#!/bin/bash
mkdir -p gpurun_out
QR_TRACE_PASSES=1 timeout 600 python scripts/prof_run.py --n 30 --L 6 --reps 2 --seed 1234 --opt axis_plan=1 > gpurun_out/axis3_trace.log 2>&1
timeout 900 python scripts/ab_axis_plan.py --cases 30x30,28x30,20x20 --modes 0,1 --tile-bits 0 --reps 2 > gpurun_out/axis3_ab.log 2>&1
grep "nv 2" gpurun_out/axis3_trace.log | tail -18; cat gpurun_out/axis3_ab.log
