#!/bin/bash
# round 2, final source: --set full captures of one backward layer (contiguous pass + two axis-aware passes) and of a
# three-round contiguous pass
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_tile12 -s 18 -c 6 -o gpurun_out/r2_prof_final_bwd_n30 \
    python scripts/prof_run.py --n 30 --L 6 --seed 1234 > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out/r2_prof_final_bwd_n30.ncu-rep
