#!/bin/bash
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
{
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus $NG 2>&1 | grep -E "^\{|Error|error" | cut -c1-330
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29559 bench.py --gpus $NG --workload batch14 2>&1 | grep -E "^\{|Error|error" | cut -c1-330
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29555 scripts/shard_run.py --qubits 28 --layers 3 --reps 2 --check-single 2>&1 | grep -E "^\{|PARITY|SINGLE|Error|error" | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29560 bench.py --gpus $NG --workload shard --steps 1 --warmup 3 2>&1 | grep -E "^\{|Error|error" | cut -c1-900
} 2>&1 | tee gpurun_out/scale4.log
