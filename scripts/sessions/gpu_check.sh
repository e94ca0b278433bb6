#!/bin/bash
# End-of-round GPU session: test tier, smoke, bench lines (all workloads + reference arm), ncu launch lists.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
cat gpurun_out/bench_default.json | cut -c1-600
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>&1
cut -c1-300 gpurun_out/bench_reference.json
for w in qaoa26 batch14 mcclean26 mcclean30; do
timeout 900 python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
cut -c1-400 gpurun_out/bench_$w.json
done
# PDL on/off over the register sizes around the auto threshold (L = 8 gradients, device time)
for n in 20 22; do
for pdl in 0 2; do
echo "n=$n pdl=$pdl $(timeout 120 python scripts/prof_run.py --n $n --L 8 --reps 6 --opt pdl=$pdl | tail -1)" >> gpurun_out/pdl_sweep.log
done
done
cat gpurun_out/pdl_sweep.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_n30_final.csv \
    python scripts/prof_run.py --n 30 --L 3 --tile-bits 0 > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench20_final.csv \
    python bench.py --steps 1 --warmup 1 --hbm-target 0 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
# ncu --set full of the backward tile pass: default workload (n = 20, half-size tiles) and the HBM-bound size (n = 30)
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_tile12ILi2 -s 4 -c 2 -o gpurun_out/prof_bwd_n20_v5 \
    python scripts/prof_run.py --n 20 --L 4 > gpurun_out/ncu_full_n20.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_tile12ILi2 -s 3 -c 3 -o gpurun_out/prof_bwd_n30_v5 \
    python scripts/prof_run.py --n 30 --L 2 > gpurun_out/ncu_full_n30.log 2>&1
ls -la gpurun_out/*.ncu-rep
