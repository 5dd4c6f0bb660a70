#!/bin/bash
# un-permute beside the resume launches: GPU suite, then A/B of the call time
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02v2_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02v2_tests.log
for k in 1 2; do
echo "=== un-permute beside the resume launches (default)"; timeout 300 python scripts/gpu_mini.py 2>&1 | tail -1
echo "=== un-permute behind them (DCOL_UNP_BESIDE=0)"; DCOL_UNP_BESIDE=0 timeout 300 python scripts/gpu_mini.py 2>&1 | tail -1
done > gpurun_out/r02v2_ab.txt 2>&1
tail -4 gpurun_out/r02v2_tests.log; cat gpurun_out/r02v2_ab.txt
