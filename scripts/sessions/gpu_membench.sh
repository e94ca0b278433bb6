#!/bin/bash
# Bare access patterns of the tile passes (scripts/membench.cu; build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
# -o scripts/bin/membench scripts/membench.cu).  Results: profiles/r1_membench_patterns.log
mkdir -p gpurun_out
B=scripts/bin/membench
{
# contiguous / low strides / high strides, one and two vectors, without and with L2 prefetch of the next tile
for nv in 1 2; do for pf in 0 1; do
$B 30 $nv 3 4 5 6 7 8 9 10 11 $pf
$B 30 $nv 12 13 14 15 16 17 18 19 20 $pf
$B 30 $nv 21 22 23 24 25 26 27 28 29 $pf
done; done
# row width sweep (128 B ... 2 KiB rows), low and high strides
$B 30 2 3 12 13 14 15 16 17 18 19 0
$B 30 2 3 4 12 13 14 15 16 17 18 0
$B 30 2 3 4 5 12 13 14 15 16 17 0
$B 30 2 3 22 23 24 25 26 27 28 29 0
$B 30 2 3 4 23 24 25 26 27 28 29 0
# CTA pairs on adjacent tiles (cluster of 2, cluster barrier per tile), and out-of-place
for bits in "12 13 14 15 16 17 18 19 20" "21 22 23 24 25 26 27 28 29"; do
$B 30 2 $bits 2 2 148
$B 30 2 $bits 2 4 148
done
$B 30 2 3 4 5 6 7 8 9 10 11 0 1 148 1
} 2>&1 | tee gpurun_out/membench.log
