#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
nvidia-smi topo -m 2>&1 | head -8
timeout 600 python -m pytest tests/test_sharded.py -m gpu -x -q 2>&1 | tail -3
for n in 24 29; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 scripts/shard_run.py --qubits $n --layers 3 --reps 2 --check-single 2>&1 | grep -E "^\{|PARITY|SINGLE|Error|error" | head
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 scripts/shard_run.py --qubits 31 --layers 3 --reps 2 2>&1 | grep -E "^\{|PARITY|SINGLE|Error|error" | head
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | grep -E "^\{|Error|error" | head -3
