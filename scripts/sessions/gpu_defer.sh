#!/bin/bash
# Full GPU test tier (timed) + A/B of QR_OPT_DEFER_REDUCE (one reduction launch per gradient).
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -16 gpurun_out/pytest_gpu.log
run() {
  name=$1; w=$2; shift 2
  timeout 300 python bench.py --workload $w --no-cpu-baseline --hbm-target 0 "$@" > gpurun_out/df_${name}.json 2>> gpurun_out/df.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/df_${name}.json")); s=d["sched"]
    print("%-24s ms=%.4f e2e_ms=%.4f fwd=%.4f bwd=%.4f launches=%d %s" % ("${name}", d["ms_per_step"], d["e2e"]["ms_per_step"], s["ms_forward"], s["ms_backward"], d["gpu_launches"], d["clocks"]["reasons"]))
except Exception as e:
    print("${name} FAILED", e)
PY
}
run n20_defer1 mcclean20 --steps 40
run n20_defer0 mcclean20 --steps 40 --opt defer_reduce=0
run n20_defer1b mcclean20 --steps 40
run n20_defer0b mcclean20 --steps 40 --opt defer_reduce=0
run n20_defer1_pdl2 mcclean20 --steps 40 --opt pdl=2
run q26_defer1 qaoa26
run q26_defer0 qaoa26 --opt defer_reduce=0
run n26_defer1 mcclean26
run n26_defer0 mcclean26 --opt defer_reduce=0
run n30_defer1 mcclean30 --warmup 1 --steps 2
run n30_defer0 mcclean30 --warmup 1 --steps 2 --opt defer_reduce=0
tail -5 gpurun_out/df.err
