#!/bin/bash
# round 2: A/B of the axis-aware plans + launch list at n = 30
mkdir -p gpurun_out
timeout 900 python scripts/ab_axis_plan.py > gpurun_out/axis_ab.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/axis_launches_n30.csv python scripts/prof_run.py --n 30 --L 3 > gpurun_out/axis_ncu.log 2>&1
cat gpurun_out/axis_ab.log
