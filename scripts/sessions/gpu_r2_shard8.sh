#!/bin/bash
# Round 2, 8 GPUs: bench line (30x30 sharded strong scaling + parity vs one GPU + config 4 + config 5 = 33x20).
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node ${1:-8} --master-addr 127.0.0.1 --master-port 29621"
{
nvidia-smi --query-gpu=name,clocks.sm --format=csv,noheader | head -2
echo "== bench --gpus ${1:-8}"
timeout 1200 $TR bench.py --gpus ${1:-8} --steps 3 --warmup 1 2>&1 | grep -v "^\*\|OMP_NUM" | tail -3
} 2>&1 | tee gpurun_out/r2_shard${1:-8}.log
