#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_tile12ILi2ELb0ELi0ELi11ELb1 -s 1 -c 2 -o gpurun_out/prof_bwd_pair \
    python scripts/prof_run.py --n 28 --L 2 --tile-bits 12 --opt pair=1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
