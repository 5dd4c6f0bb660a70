#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in 1 2; do
echo "=== un-permute beside the resume launches, low-priority stream (default)"; timeout 300 python scripts/gpu_mini.py 2>&1 | tail -1
echo "=== un-permute behind them (DCOL_UNP_BESIDE=0)"; DCOL_UNP_BESIDE=0 timeout 300 python scripts/gpu_mini.py 2>&1 | tail -1
done > gpurun_out/r02w2_ab.txt 2>&1
cat gpurun_out/r02w2_ab.txt
