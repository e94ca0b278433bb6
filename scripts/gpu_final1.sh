#!/bin/bash
# full check of the current default configuration: GPU tests, smoke, bench (default + reference arm), launch lists, ncu captures
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
cat gpurun_out/bench_default.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>&1
cat gpurun_out/bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_n30_default.csv \
    python scripts/prof_run.py --n 30 --L 3 > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench20.csv \
    python bench.py --steps 1 --warmup 1 --hbm-target 0 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_tile12ILi2 -s 3 -c 3 -o gpurun_out/prof_bwd_final \
    python scripts/prof_run.py --n 30 --L 2 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_tile12ILi1 -s 3 -c 3 -o gpurun_out/prof_fwd_final \
    python scripts/prof_run.py --n 30 --L 2 > gpurun_out/ncu_full_fwd.log 2>&1
ls -la gpurun_out | tail -6
