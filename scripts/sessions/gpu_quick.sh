#!/bin/bash
# Quick GPU session: test tier + 20x20 bench lines (default, PDL off) + batch14 / QAOA-26.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
run() {
  name=$1; w=$2; shift 2
  timeout 300 python bench.py --workload $w --no-cpu-baseline --hbm-target 0 "$@" > gpurun_out/q_${name}.json 2>> gpurun_out/q.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/q_${name}.json")); s=d["sched"]
    print("%-24s ms=%.4f e2e_ms=%.4f fwd=%.4f bwd=%.4f launches=%d %s" % ("${name}", d["ms_per_step"], d["e2e"]["ms_per_step"], s["ms_forward"], s["ms_backward"], d["gpu_launches"], d["clocks"]["reasons"]))
except Exception as e:
    print("${name} FAILED", e)
PY
}
run n20 mcclean20 --steps 40
run n20_pdl0 mcclean20 --steps 40 --opt pdl=0
run n20_b mcclean20 --steps 40
run batch14 batch14
run qaoa26 qaoa26
for n in 14 16 18 22; do
echo "n=$n $(timeout 120 python scripts/prof_run.py --n $n --L 8 --reps 6 | tail -1)"
done
tail -5 gpurun_out/q.err
