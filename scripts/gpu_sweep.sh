#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for cfg in "--tile-bits 12 --tile-bits-x 11 --min-row-bits 2" "--tile-bits 12 --tile-bits-x 12 --min-row-bits 2" "--tile-bits 12 --tile-bits-x 11 --min-row-bits 2 --async-bwd 1" "--tile-bits 11 --tile-bits-x 11 --min-row-bits 2" "--tile-bits 12 --tile-bits-x 10 --min-row-bits 1" "--tile-bits 10 --tile-bits-x 10 --min-row-bits 1" "--tile-bits 11 --tile-bits-x 11 --min-row-bits 1"; do
  timeout 300 python scripts/prof_run.py --n 30 --L 3 --reps 2 $cfg 2>&1 | tail -2
done
