#!/bin/bash
# Round 2, 2 GPUs: swap engine vs peer engine on a 30-qubit register, parity against the one-GPU path, bench line.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
{
nvidia-smi --query-gpu=name,clocks.sm --format=csv,noheader
timeout 900 python -m pytest tests/test_sharded.py -q -m gpu -x 2>&1 | tail -3
echo "== swap 30x4"
timeout 600 $TR scripts/shard_run.py --qubits 30 --layers 4 --reps 3 --check-single --mode swap 2>&1 | grep -v "^\*\|OMP_NUM" | tail -4
echo "== peer 30x4"
timeout 600 $TR scripts/shard_run.py --qubits 30 --layers 4 --reps 3 --mode peer 2>&1 | grep -v "^\*\|OMP_NUM" | tail -3
echo "== swap 31x4"
timeout 600 $TR scripts/shard_run.py --qubits 31 --layers 4 --reps 2 --mode swap 2>&1 | grep -v "^\*\|OMP_NUM" | tail -3
echo "== bench --gpus 2"
timeout 900 $TR bench.py --gpus 2 --steps 2 --warmup 1 2>&1 | grep -v "^\*\|OMP_NUM" | tail -3
} 2>&1 | tee gpurun_out/r2_shard2.log
