#!/bin/bash
# full GPU test tier + smoke on one GPU
mkdir -p gpurun_out
{
timeout 2400 python -m pytest tests -q -m gpu -x --durations=8 2>&1 | tail -18
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
} 2>&1 | tee gpurun_out/r2_full1.log
