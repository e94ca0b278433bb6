#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -q -d POWER,CLOCK | head -60 > gpurun_out/diag_smi.txt
for n in 28 30; do
timeout 300 python scripts/diag_clocks.py --n $n --L 6 --opt lean=3 > gpurun_out/diag_n${n}_lean.log 2>&1
timeout 300 python scripts/diag_clocks.py --n $n --L 6 --opt lean=0 > gpurun_out/diag_n${n}_old.log 2>&1
done
grep -h "^n=" gpurun_out/diag_n*.log
