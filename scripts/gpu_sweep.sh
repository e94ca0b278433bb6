#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for cfg in "--prefetch 0" "--prefetch 1" "--async-bwd 1" "--async-bwd 1 --async-fwd 1" "--tile-bits 11 --prefetch 1"; do
timeout 300 python scripts/prof_run.py --n 30 --L 3 --reps 2 $cfg 2>&1 | tail -1
done
