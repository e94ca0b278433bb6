#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q -k "batch" 2>&1 | tail -2
for mb in 256 64 32 16 1024; do
echo "chunk $mb"
timeout 600 python bench.py --workload batch14 --no-cpu-baseline --batch-chunk-mb $mb --steps 2 --warmup 1 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.0f e2e %.0f ms/step %.1f e2e_ms %.1f launches %d'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['e2e']['ms_per_step'],d['gpu_launches']))"
done
echo "tile 11"; timeout 600 python bench.py --workload batch14 --no-cpu-baseline --tile-bits 11 --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-200
