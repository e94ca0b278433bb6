#!/bin/bash
# ncu full captures of the backward and forward tile passes of one layer at n=30 (16 GiB vectors, HBM-bound);
# summaries for profiles/: python scripts/ncu_summary.py gpurun_out/prof_bwd.ncu-rep profiles/<name>.csv
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_tile12ILi2 -s 3 -c 3 -o gpurun_out/prof_bwd \
    python scripts/prof_run.py --n 30 --L 2 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_tile12ILi1 -s 3 -c 3 -o gpurun_out/prof_fwd \
    python scripts/prof_run.py --n 30 --L 2 > gpurun_out/ncu_full_fwd.log 2>&1
ls -la gpurun_out | tail -4
