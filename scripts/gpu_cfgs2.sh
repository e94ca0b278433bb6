#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/cfgs2.log; rm -f $L
b() { timeout 300 python bench.py --no-cpu-baseline --hbm-target 0 "$@" 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(sys.argv[1:], d['config']['workload'][:30], 'value=%.4g' % d['value'], 'ms=%.4g' % d['ms_per_step'], 'fwd=%.4g bwd=%.4g' % (d['sched']['ms_forward'], d['sched']['ms_backward']))
" "$@" >> $L; }
b --workload mcclean20 --opt staged=1
b --workload mcclean20 --opt staged=3
b --workload mcclean20 --opt staged=2
b --workload mcclean26 --opt staged_min_bit=19
b --workload mcclean26 --opt staged_min_bit=12
b --workload mcclean26 --opt staged=1
b --workload mcclean26 --opt staged=0
b --workload mcclean26 --opt prefetch=5
b --workload qaoa26 --opt staged_min_bit=19
b --workload qaoa26 --opt staged=1
b --workload batch14 --opt staged=1
b --workload batch14 --opt prefetch=0
cat $L
