#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
NG=$(nvidia-smi -L | wc -l)
echo "GPUS=$NG"
{
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29555 scripts/shard_run.py --qubits 29 --layers 3 --reps 2 --check-single 2>&1 | grep -E "^\{|PARITY|SINGLE|Error|error" | head
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29556 scripts/shard_run.py --qubits 33 --layers 2 --reps 2 2>&1 | grep -E "^\{|PARITY|SINGLE|Error|error" | head
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29558 scripts/shard_run.py --qubits 33 --layers 20 --reps 1 2>&1 | grep -E "^\{|PARITY|SINGLE|Error|error" | head
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus $NG --steps 5 --warmup 3 2>&1 | grep -E "^\{|Error|error" | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29559 bench.py --gpus $NG --workload batch14 --steps 3 --warmup 3 2>&1 | grep -E "^\{|Error|error" | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29560 bench.py --gpus $NG --workload shard --steps 2 --warmup 3 2>&1 | grep -E "^\{|Error|error" | cut -c1-1200
} 2>&1 | tee gpurun_out/shard8.log
