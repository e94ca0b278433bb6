#!/bin/bash
# Sharded path with / without the exchange-free Rz on global qubits (QR_OPT_SHARD_ZSKIP); run with gpurun --gpus N.
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "GPUS=$NG"
NQ=$((28 + $(python -c "import math;print(int(math.log2($NG)))")))
run() { port=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $port scripts/shard_run.py "$@" 2>&1 | grep -E "^\{|PARITY|SINGLE|Error|error" | cut -c1-600; }
{
run 29555 --qubits $NQ --layers 4 --reps 2 --check-single
run 29556 --qubits $NQ --layers 4 --reps 2 --check-single --opt shard_zskip=0
if [ "$NG" = "8" ]; then
run 29557 --qubits 33 --layers 20 --reps 1
run 29558 --qubits 33 --layers 20 --reps 1 --opt shard_zskip=0
fi
} 2>&1 | tee gpurun_out/shardz_$NG.log
