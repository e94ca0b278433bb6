#!/bin/bash
for i in 1 2; do
timeout 200 python scripts/diag_clocks.py --n 30 --L 3 --opt tile_bits=12 2>&1 | grep "^n=" | tail -1
QRADIENT_B200_LIB=$PWD/scripts/bin/libqr_prev.so timeout 200 python scripts/diag_clocks.py --n 30 --L 3 --opt tile_bits=12 2>&1 | grep "^n=" | tail -1
done
