#!/bin/bash
# Round 2: self-staged loads (QR_OPT_SELF_STAGE) vs the round-1 default, n = 30 / 26 / 20; parity of the new variant first.
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_parity.py -q -m gpu -k "kernel_variants and cuda" -x 2>&1 | tail -3
for ss in 0 1 5 3 7; do
  echo "== mcclean30 self_stage=$ss"
  timeout 300 python bench.py --workload mcclean30 --steps 2 --warmup 1 --no-cpu-baseline --opt self_stage=$ss 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['ms_per_step'], d['sched']['ms_forward'], d['sched']['ms_backward'], d['roofline']['ms_per_launch'], d['sched']['frac_of_peak'])
    else: print(l.rstrip())
"
done
for ss in 0 3 15; do
  echo "== mcclean26 self_stage=$ss"
  timeout 300 python bench.py --workload mcclean26 --steps 5 --warmup 2 --no-cpu-baseline --opt self_stage=$ss 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['ms_per_step'], d['sched']['ms_forward'], d['sched']['ms_backward'], d['roofline']['ms_per_launch'], d['sched']['frac_of_peak'])
    else: print(l.rstrip())
"
  echo "== mcclean20 self_stage=$ss"
  timeout 300 python bench.py --workload mcclean20 --steps 20 --warmup 3 --no-cpu-baseline --hbm-target 0 --opt self_stage=$ss 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['ms_per_step'], d['sched']['ms_forward'], d['sched']['ms_backward'], d['roofline']['ms_per_launch'], d['e2e']['ms_per_step'])
    else: print(l.rstrip())
"
done
} 2>&1 | tee gpurun_out/r2_selfstage.log
