#!/bin/bash
# A/B of the split rounds (QR_T12_SPLIT_XCHG): default library vs a build with -DQR_T12_SPLIT_XCHG=0; doubles as validation run.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
run() {
  name=$1; lib=$2; w=$3; shift 3
  QRADIENT_B200_LIB=$lib timeout 300 python bench.py --workload $w --no-cpu-baseline --hbm-target 0 "$@" > gpurun_out/s_${name}.json 2>> gpurun_out/s.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s_${name}.json")); s=d["sched"]
    print("%-20s ms=%.4f e2e_ms=%.4f fwd=%.4f bwd=%.4f frac=%.3f %s" % ("${name}", d["ms_per_step"], d["e2e"]["ms_per_step"], s["ms_forward"], s["ms_backward"], s["frac_of_peak"], d["clocks"]["reasons"]))
except Exception as e:
    print("${name} FAILED", e)
PY
}
NEW=$PWD/qradient_b200/libqradient_b200.so
OLD=$PWD/qradient_b200/libqradient_b200_nosplit.so
run n20_new $NEW mcclean20 --steps 40
run n20_old $OLD mcclean20 --steps 40
run n30_new $NEW mcclean30 --warmup 1 --steps 2
run n30_old $OLD mcclean30 --warmup 1 --steps 2
run n26_new $NEW mcclean26 --steps 5
run n26_old $OLD mcclean26 --steps 5
run q26_new $NEW qaoa26
run b14_new $NEW batch14
timeout 600 python bench.py > gpurun_out/bench_default_split.json 2> gpurun_out/bench_default_split.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/bench_default_split.json
tail -3 gpurun_out/s.err
