#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --workload qaoa26 --no-cpu-baseline --steps 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('qaoa26 value %.2f e2e %.2f ms/step %.1f'%(d['value'],d['e2e']['value'],d['ms_per_step']), d['roofline']['achieved'], d['sched'])"
timeout 600 python bench.py --workload batch14 --no-cpu-baseline --steps 2 --warmup 1 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('batch14 value %.0f e2e %.0f ms/step %.1f e2e_ms %.1f launches %d'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['e2e']['ms_per_step'],d['gpu_launches']))"
timeout 600 python bench.py --workload batch14 --no-cpu-baseline --steps 2 --warmup 1 --batch-chunk-mb 2048 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('batch14/2048 value %.0f e2e %.0f ms/step %.1f e2e_ms %.1f launches %d'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['e2e']['ms_per_step'],d['gpu_launches']))"
