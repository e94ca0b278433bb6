#!/bin/bash
# per-kind device times of the swap engine (rank 0): N given as $1; slices / SM split variants
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631"
{
for opts in "shard_slices=1" "shard_slices=8" "shard_slices=8 --opt shard_xsms=48" "shard_slices=4 --opt shard_xsms=56" "shard_slices=2 --opt shard_xsms=56"; do
echo "== $opts"
timeout 600 $TR scripts/shard_run.py --qubits 30 --layers 6 --reps 3 --mode swap --opt $opts 2>&1 | grep "^{\|SINGLE"
done
} 2>&1 | tee gpurun_out/r2_shardt_$N.log
