#!/bin/bash
mkdir -p gpurun_out
B=scripts/bin/membench
{
echo "# row width sweep, low strides (bits from 12 up)"
$B 30 2 12 13 14 15 16 17 18 19 20 0
$B 30 2 3 12 13 14 15 16 17 18 19 0
$B 30 2 3 4 12 13 14 15 16 17 18 0
$B 30 2 3 4 5 12 13 14 15 16 17 0
$B 30 2 3 4 5 6 12 13 14 15 16 0
echo "# row width sweep, high strides (bits from 29 down)"
$B 30 2 21 22 23 24 25 26 27 28 29 0
$B 30 2 3 22 23 24 25 26 27 28 29 0
$B 30 2 3 4 23 24 25 26 27 28 29 0
$B 30 2 3 4 5 24 25 26 27 28 29 0
$B 30 2 3 4 5 6 25 26 27 28 29 0
echo "# nv=1 row width sweep high strides"
$B 30 1 21 22 23 24 25 26 27 28 29 0
$B 30 1 3 22 23 24 25 26 27 28 29 0
$B 30 1 3 4 23 24 25 26 27 28 29 0
$B 30 1 3 4 5 24 25 26 27 28 29 0
echo "# mid strides"
$B 30 2 15 16 17 18 19 20 21 22 23 0
$B 30 2 18 19 20 21 22 23 24 25 26 0
echo "# single high bit"
for hb in 20 21 23 25 27 29; do
$B 30 2 12 13 14 15 16 17 18 19 $hb 0
done
} 2>&1 | tee gpurun_out/membench2.log
