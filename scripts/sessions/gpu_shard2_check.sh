#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded.py -m gpu -x -q 2>&1 | tail -2
for n in 24 27; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 scripts/shard_run.py --qubits $n --layers 3 --reps 2 --tile-bits 0 --check-single 2>&1 | grep -E "^\{|PARITY|SINGLE|Error|error" | cut -c1-300
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus 2 2>&1 | grep -E "^\{|Error|error" | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29558 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 2>&1 | grep -E "^\{|Error|error" | cut -c1-200
