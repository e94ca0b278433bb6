#!/bin/bash
mkdir -p gpurun_out
B=scripts/bin/membench
{
for bits in "12 13 14 15 16 17 18 19 20" "21 22 23 24 25 26 27 28 29"; do
$B 30 2 $bits 0 1 148
$B 30 2 $bits 0 2 148
$B 30 2 $bits 2 2 148
$B 30 2 $bits 2 4 148
$B 30 2 $bits 0 4 148
$B 30 2 $bits 2 8 144
$B 30 2 $bits 3 1 148
$B 30 2 $bits 0 1 296
$B 30 1 $bits 2 2 148
$B 30 1 $bits 2 4 148
done
} 2>&1 | tee gpurun_out/membench3.log
