#!/bin/bash
# Round 2: L2 prefetch-size hints (ld.global.L2::256B -> LDG.E.LTC256B) on the strided 128 B-row patterns.
mkdir -p gpurun_out
B=scripts/bin/membench
LOW="12 13 14 15 16 17 18 19 20"
HIGH="21 22 23 24 25 26 27 28 29"
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for nv in 2 1; do
for bits in "$LOW" "$HIGH"; do
for mode in 0 4 5 6 7 3; do
$B 30 $nv $bits $mode 1 148
done
# L2 fetch granularity limit 128 B with plain loads, and with the hint
$B 30 $nv $bits 0 1 148 0 128
$B 30 $nv $bits 4 1 148 0 128
# out of place with the hint
$B 30 $nv $bits 4 1 148 1
# two CTAs per SM worth of grid
$B 30 $nv $bits 4 1 296
done
done
# contiguous reference
$B 30 2 3 4 5 6 7 8 9 10 11 0 1 148
$B 30 2 3 4 5 6 7 8 9 10 11 4 1 148
} 2>&1 | tee gpurun_out/membench2.log
