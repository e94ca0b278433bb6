#!/bin/bash
# bulk asynchronous copies from the peer (cp.async.bulk + mbarrier) against plain SM loads and the copy engines
mkdir -p gpurun_out
B=scripts/bin/p2pbench
{
timeout 60 $B 0 148 8 4096; timeout 60 $B 6 148 8 4096
for kib in 4 8 16 32; do for ctas in 148 296 592; do timeout 60 $B 7 $ctas $kib 4096; done; done
} 2>&1 | tee gpurun_out/r2_p2pbench_bulk.log
