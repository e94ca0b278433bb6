#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/diag_ch.log
run() { timeout 120 python scripts/diag_clocks.py --n 30 --L 3 "$@" 2>&1 | grep "^n=" | tail -1 >> gpurun_out/diag_ch.log; }
run --opt cache_hints=0
run --opt cache_hints=1
run --opt cache_hints=2
run --opt cache_hints=3
run --opt cache_hints=4
run --opt cache_hints=12
cat gpurun_out/diag_ch.log
