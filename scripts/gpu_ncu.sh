#!/bin/bash
# ncu full captures of the backward and forward tile passes (n=28: 4 GiB vectors, HBM-bound)
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_tile_passILi2ELi3 -s 1 -c 2 -o gpurun_out/prof_bwd_r1 \
    python scripts/prof_run.py --n 28 --L 3 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_tile_passILi1ELi3 -s 1 -c 2 -o gpurun_out/prof_fwd_r1 \
    python scripts/prof_run.py --n 28 --L 3 > gpurun_out/ncu_full_fwd.log 2>&1
tail -3 gpurun_out/ncu_full_fwd.log
ls -la gpurun_out
