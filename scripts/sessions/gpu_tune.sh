#!/bin/bash
# Cheap option A/Bs on the short-pass and batched workloads.
mkdir -p gpurun_out
run() {
  name=$1; w=$2; shift 2
  timeout 300 python bench.py --workload $w --no-cpu-baseline --hbm-target 0 "$@" > gpurun_out/t_${name}.json 2>> gpurun_out/t.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/t_${name}.json")); s=d["sched"]
    print("%-24s ms=%.4f e2e_ms=%.4f fwd=%.4f bwd=%.4f launches=%d %s" % ("${name}", d["ms_per_step"], d["e2e"]["ms_per_step"], s["ms_forward"], s["ms_backward"], d["gpu_launches"], d["clocks"]["reasons"]))
except Exception as e:
    print("${name} FAILED", e)
PY
}
run n20 mcclean20 --steps 40
run n20_pf0 mcclean20 --steps 40 --prefetch 0
run n20_minrow3 mcclean20 --steps 40 --opt min_row_bits=3 --tile-bits 11
run n20_fwd2 mcclean20 --steps 40 --ctas-fwd 1
run b14 batch14
run b14_c128 batch14 --batch-chunk-mb 128
run b14_c512 batch14 --batch-chunk-mb 512
run b14_c1024 batch14 --batch-chunk-mb 1024
run b14_pdl0 batch14 --opt pdl=0
tail -3 gpurun_out/t.err
