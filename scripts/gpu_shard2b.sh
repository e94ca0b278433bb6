#!/bin/bash
mkdir -p gpurun_out
for n in 30 31; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 scripts/shard_run.py --qubits $n --layers 4 --reps 2 2>&1 | grep -E "^\{|PARITY|SINGLE|Error|error" | head
done
