#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/diag_skew.log
for skew in 0 256 4352 69888 1118464 3215616 34672896; do
timeout 300 python scripts/diag_clocks.py --n 30 --L 3 --opt buf_skew=$skew --opt lean=3 2>&1 | grep "^n=" | tail -1 >> gpurun_out/diag_skew.log
done
timeout 300 python scripts/diag_clocks.py --n 30 --L 3 --opt buf_skew=69888 --opt lean=0 2>&1 | grep "^n=" | tail -1 >> gpurun_out/diag_skew.log
timeout 300 python scripts/diag_clocks.py --n 29 --L 3 --opt buf_skew=0 --opt lean=3 2>&1 | grep "^n=" | tail -1 >> gpurun_out/diag_skew.log
timeout 300 python scripts/diag_clocks.py --n 29 --L 3 --opt buf_skew=69888 --opt lean=3 2>&1 | grep "^n=" | tail -1 >> gpurun_out/diag_skew.log
cat gpurun_out/diag_skew.log
