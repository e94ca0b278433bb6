#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
L=gpurun_out/small.log; rm -f $L
b() { timeout 300 python bench.py --no-cpu-baseline --hbm-target 0 "$@" 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(sys.argv[1:], d['config']['workload'][:30], 'value=%.4g' % d['value'], 'e2e=%.4g' % d['e2e']['value'], 'ms=%.4g' % d['ms_per_step'], 'fwd=%.4g bwd=%.4g' % (d['sched']['ms_forward'], d['sched']['ms_backward']))
" "$@" >> $L; }
b --workload mcclean20
b --workload mcclean20 --opt prefetch=0
b --workload mcclean20 --opt tile_bits=11 --opt min_row_bits=2
b --workload mcclean3
b --workload batch14
b --workload qaoa26
cat $L
timeout 200 python scripts/diag_clocks.py --n 30 --L 3 2>&1 | grep "^n=" | tail -1
timeout 200 python scripts/diag_clocks.py --n 16 --L 16 2>&1 | grep "^n=" | tail -1
timeout 200 python scripts/diag_clocks.py --n 12 --L 12 2>&1 | grep "^n=" | tail -1
