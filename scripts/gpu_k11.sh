#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
L=gpurun_out/k11.log; rm -f $L
b() { timeout 300 python bench.py --no-cpu-baseline --hbm-target 0 "$@" 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(sys.argv[1:], d['config']['workload'][:30], 'value=%.4g' % d['value'], 'ms=%.4g' % d['ms_per_step'], 'fwd=%.4g bwd=%.4g' % (d['sched']['ms_forward'], d['sched']['ms_backward']), 'P', d['config']['passes_per_layer'])
" "$@" >> $L; }
b --workload mcclean20
b --workload mcclean20 --opt tile_bits=11 --opt min_row_bits=2
b --workload mcclean20 --opt tile_bits=11
b --workload mcclean26
b --workload mcclean26 --opt tile_bits=11
b --workload qaoa26
b --workload qaoa26 --opt tile_bits=11
b --workload batch14
b --workload batch14 --opt tile_bits=11
cat $L
for n in 22 24 27 28; do
timeout 200 python scripts/diag_clocks.py --n $n --L 4 2>&1 | grep "^n=" | tail -1
timeout 200 python scripts/diag_clocks.py --n $n --L 4 --opt tile_bits=11 2>&1 | grep "^n=" | tail -1
done
