#!/bin/bash
for cfg in "--prefetch 1" "--prefetch 2" "--prefetch 3" "--prefetch 5"; do
timeout 300 python scripts/prof_run.py --n 30 --L 3 --reps 2 $cfg 2>&1 | tail -1
done
timeout 300 python scripts/prof_run.py --n 20 --L 20 --reps 5 --prefetch 0 2>&1 | tail -1
timeout 300 python scripts/prof_run.py --n 20 --L 20 --reps 5 --prefetch 1 2>&1 | tail -1
