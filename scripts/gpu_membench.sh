#!/bin/bash
mkdir -p gpurun_out
B=scripts/bin/membench
{
for nv in 1 2; do
for pf in 0 1; do
$B 30 $nv 3 4 5 6 7 8 9 10 11 $pf          # contiguous (like pass 0)
$B 30 $nv 12 13 14 15 16 17 18 19 20 $pf   # old pass 1
$B 30 $nv 21 22 23 24 25 26 27 28 29 $pf   # old pass 2
$B 30 $nv 12 13 17 18 19 20 21 22 23 $pf   # new pass 1
$B 30 $nv 14 15 16 24 25 26 27 28 29 $pf   # new pass 2
done
done
# single high bits added to a low set
for hb in 17 19 21 23 25 27 29; do
$B 30 2 12 13 14 15 16 17 18 19 $hb 1
done
# which local position holds the high bits: lanes (local 3,4) vs registers (local 9-11)
$B 30 2 27 28 29 12 13 14 15 16 17 1
$B 30 2 12 13 14 15 16 17 27 28 29 1
$B 30 2 12 13 14 27 28 29 15 16 17 1
# n = 28 for reference
$B 28 2 12 13 14 15 16 17 18 19 20 1
$B 28 2 19 20 21 22 23 24 25 26 27 1
} 2>&1 | tee gpurun_out/membench.log
