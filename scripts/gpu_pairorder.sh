#!/bin/bash
# A/B of QR_OPT_PAIR_ORDER (adjacent tile pairs per CTA in the strided passes) at the HBM-bound sizes.
mkdir -p gpurun_out
run() {  # name, workload, extra args
  name=$1; w=$2; shift 2
  timeout 300 python bench.py --workload $w --no-cpu-baseline --warmup 1 --steps 2 "$@" > gpurun_out/po_${name}.json 2>> gpurun_out/po.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/po_${name}.json")); s=d["sched"]
    print("%-28s ms=%.1f fwd=%.1f bwd=%.1f frac=%.3f %s" % ("${name}", d["ms_per_step"], s["ms_forward"], s["ms_backward"], s["frac_of_peak"], d["clocks"]["reasons"]))
except Exception as e:
    print("${name} FAILED", e)
PY
}
timeout 600 python -m pytest tests/test_parity.py -m gpu -x -q -k "pair_order" > gpurun_out/pytest_po.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_po.log
tail -3 gpurun_out/pytest_po.log
run n30_default mcclean30
run n30_po1 mcclean30 --opt pair_order=1
run n30_po1_nostage mcclean30 --opt pair_order=1 --opt staged=0
run n30_po2_pf mcclean30 --opt pair_order=2 --opt prefetch=5
run n30_po2 mcclean30 --opt pair_order=2
run n30_po7 mcclean30 --opt pair_order=7
run n30_po7_nostage mcclean30 --opt pair_order=7 --opt staged=0
run n30_default2 mcclean30
run n26_default mcclean26
run n26_po1 mcclean26 --opt pair_order=1
run n26_po7 mcclean26 --opt pair_order=7
run n26_k12_po7 mcclean26 --opt pair_order=7 --tile-bits 12
tail -5 gpurun_out/po.err
