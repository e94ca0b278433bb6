#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python scripts/prof_run.py --n 20 --L 20 --reps 5 2>&1 | tail -1
timeout 300 python scripts/prof_run.py --n 30 --L 3 --reps 2 2>&1 | tail -1
timeout 300 python scripts/prof_run.py --n 30 --L 3 --reps 2 --tile-bits 11 2>&1 | tail -1
timeout 300 python scripts/prof_run.py --n 26 --L 5 --reps 2 2>&1 | tail -1
