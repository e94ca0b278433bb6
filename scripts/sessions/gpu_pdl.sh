#!/bin/bash
# A/B of programmatic dependent launch (QR_OPT_PDL) for the tile passes: correctness first, then bench lines.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
timeout 600 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "programmatic or config2" > gpurun_out/pytest_pdl.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_pdl.log
tail -5 gpurun_out/pytest_pdl.log
for pdl in 0 2 0 2; do
timeout 300 python bench.py --hbm-target 0 --no-cpu-baseline --steps 40 --opt pdl=$pdl > gpurun_out/bench20_pdl${pdl}.json 2>> gpurun_out/bench_pdl.err
python - <<PY
import json; d=json.load(open("gpurun_out/bench20_pdl${pdl}.json")); print("n20 pdl=${pdl}", d["ms_per_step"], d["e2e"]["ms_per_step"], d["sched"], d["roofline"]["ms_per_launch"], d["clocks"])
PY
done
for w in batch14 qaoa26 mcclean26; do
for pdl in 0 2; do
timeout 300 python bench.py --workload $w --no-cpu-baseline --opt pdl=$pdl > gpurun_out/bench_${w}_pdl${pdl}.json 2>> gpurun_out/bench_pdl.err
python - <<PY
import json; d=json.load(open("gpurun_out/bench_${w}_pdl${pdl}.json")); print("${w} pdl=${pdl}", d["ms_per_step"], d["value"], d["sched"])
PY
done
done
for pdl in 0 2; do
timeout 300 python bench.py --workload mcclean30 --no-cpu-baseline --warmup 1 --steps 2 --opt pdl=$pdl > gpurun_out/bench_mcclean30_pdl${pdl}.json 2>> gpurun_out/bench_pdl.err
python - <<PY
import json; d=json.load(open("gpurun_out/bench_mcclean30_pdl${pdl}.json")); print("n30 pdl=${pdl}", d["ms_per_step"], d["sched"], d["clocks"])
PY
done
tail -5 gpurun_out/bench_pdl.err
timeout 300 python bench.py --hbm-target 0 --no-cpu-baseline --steps 40 --tile-bits 12 --opt pdl=2 > gpurun_out/bench20_k12_pdl2.json 2>> gpurun_out/bench_pdl.err
python - <<PY
import json; d=json.load(open("gpurun_out/bench20_k12_pdl2.json")); print("n20 k12 pdl=2", d["ms_per_step"], d["sched"])
PY
