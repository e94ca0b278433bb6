#!/bin/bash
mkdir -p gpurun_out
B=scripts/bin/membench
{
for nv in 1 2; do
$B 30 $nv 3 4 5 6 7 8 9 10 11 0 1 148 0
$B 30 $nv 3 4 5 6 7 8 9 10 11 0 1 148 1
$B 30 $nv 3 4 5 6 7 8 9 10 11 0 1 296 1
$B 30 $nv 12 13 14 15 16 17 18 19 20 0 1 148 1
done
$B 29 2 3 4 5 6 7 8 9 10 11 0 1 148 0
$B 29 2 3 4 5 6 7 8 9 10 11 0 1 148 1
} 2>&1 | tee gpurun_out/membench4.log
