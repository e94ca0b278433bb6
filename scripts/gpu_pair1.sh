#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/pair_check.py > gpurun_out/pair_check.log 2>&1; echo "rc=$?" >> gpurun_out/pair_check.log
tail -40 gpurun_out/pair_check.log
for opt in "pair=0" "pair=1" "pair=3" "pair=1 --opt prefetch=0" "pair=3 --opt prefetch=0"; do
timeout 120 python scripts/diag_clocks.py --n 30 --L 3 --opt tile_bits=12 --opt $opt 2>&1 | grep "^n=" | tail -1
done
