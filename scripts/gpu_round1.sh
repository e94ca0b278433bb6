#!/bin/bash
# First GPU session: parity tests, smoke, bench, ncu launch list + one full capture.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
cat gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
for pf in 0 1; do
  timeout 300 python scripts/prof_run.py --n 30 --L 3 --reps 2 --prefetch $pf >> gpurun_out/sweep30.log 2>&1
done
timeout 300 python scripts/prof_run.py --n 30 --L 3 --reps 2 --prefetch 1 --ctas-fwd 3 >> gpurun_out/sweep30.log 2>&1
timeout 300 python scripts/prof_run.py --n 30 --L 3 --reps 2 --prefetch 0 --ctas-fwd 1 >> gpurun_out/sweep30.log 2>&1
cat gpurun_out/sweep30.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 1 --warmup 1 --hbm-target 0 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_passILi2 -s 3 -c 3 -o gpurun_out/prof_bwd_r1 \
    python scripts/prof_run.py --n 28 --L 3 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_passILi1 -s 3 -c 3 -o gpurun_out/prof_fwd_r1 \
    python scripts/prof_run.py --n 28 --L 3 > gpurun_out/ncu_full_fwd.log 2>&1
ls -la gpurun_out
