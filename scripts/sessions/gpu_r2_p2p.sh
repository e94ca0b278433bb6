#!/bin/bash
mkdir -p gpurun_out
B=scripts/bin/p2pbench
{
for mode in 0 1 2 3 4 5 6; do $B $mode 148 8 4096; done
for ctas in 48 296 592; do $B 0 $ctas 8 4096; $B 4 $ctas 8 4096; done
$B 0 148 16 4096; $B 0 148 4 4096; $B 1 296 16 4096; $B 5 296 16 4096
} 2>&1 | tee gpurun_out/r2_p2pbench.log
