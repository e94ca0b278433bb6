#!/bin/bash
# lean tile kernel: parity on the GPU, n=30 timing against the generic kernel, ncu captures
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
rm -f gpurun_out/sweep30_lean.log
for lean in 3 0; do
  for pf in 1 0; do
    timeout 300 python scripts/prof_run.py --n 30 --L 3 --reps 2 --prefetch $pf --lean $lean >> gpurun_out/sweep30_lean.log 2>&1
  done
done
timeout 300 python scripts/prof_run.py --n 30 --L 3 --reps 2 --prefetch 5 --lean 3 >> gpurun_out/sweep30_lean.log 2>&1
timeout 300 python scripts/prof_run.py --n 20 --L 20 --reps 5 --prefetch 1 --lean 3 >> gpurun_out/sweep30_lean.log 2>&1
timeout 300 python scripts/prof_run.py --n 20 --L 20 --reps 5 --prefetch 1 --lean 0 >> gpurun_out/sweep30_lean.log 2>&1
cat gpurun_out/sweep30_lean.log
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_tile12ILi2 -s 1 -c 2 -o gpurun_out/prof_bwd_lean1 \
    python scripts/prof_run.py --n 28 --L 3 --prefetch 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_tile12ILi1 -s 1 -c 2 -o gpurun_out/prof_fwd_lean1 \
    python scripts/prof_run.py --n 28 --L 3 --prefetch 1 > gpurun_out/ncu_full_fwd.log 2>&1
tail -2 gpurun_out/ncu_full_fwd.log
ls -la gpurun_out | tail -8
