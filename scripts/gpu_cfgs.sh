#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/cfgs.log; rm -f $L
b() { timeout 300 python bench.py --no-cpu-baseline --hbm-target 0 "$@" 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(sys.argv[1:], d['config']['workload'][:40], 'value=%.4g' % d['value'], 'e2e=%.4g' % d['e2e']['value'], 'ms=%.4g' % d['ms_per_step'], 'bwd_ms=%.4g' % d['roofline']['ms_per_launch'], 'launches', d['gpu_launches'])
" "$@" >> $L; }
b --workload mcclean20
b --workload mcclean20 --opt lean=0
b --workload mcclean20 --opt lean=0 --opt tile_bits=11 --opt min_row_bits=2 --opt ctas_per_sm_bwd=2
b --workload mcclean20 --opt lean=0 --opt tile_bits=11 --opt min_row_bits=2 --opt ctas_per_sm_bwd=2 --opt ctas_per_sm_fwd=4
b --workload mcclean20 --opt prefetch=0
b --workload qaoa26
b --workload qaoa26 --opt lean=0
b --workload batch14
b --workload batch14 --opt lean=0
b --workload mcclean26
b --workload mcclean26 --opt lean=0
cat $L
