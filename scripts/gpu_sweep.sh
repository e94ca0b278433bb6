#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_tile_pass_dcILi2 -s 0 -c 2 -o gpurun_out/prof_bwd_dc \
    python scripts/prof_run.py --n 28 --L 3 --decoupled 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
