#!/bin/bash
# Round 2, 1 GPU: new parity tests at BASELINE sizes, swap engine on virtual shards, full bench line.
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests/test_sharded.py tests/test_gpu_scale.py -q -m gpu -x 2>&1 | tail -5
echo "== bench N=1"
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -2
} 2>&1 | tee gpurun_out/r2_check1.log
