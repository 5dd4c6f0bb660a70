#!/bin/bash
# tier launch order A/B: heavy tier first (new default) against light tier first (DCOL_TIER_ORDER=0)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in 1 2; do
echo "=== heavy tier first (default)"; timeout 300 python scripts/gpu_mini.py 2>&1 | tail -1
echo "=== light tier first (DCOL_TIER_ORDER=0)"; DCOL_TIER_ORDER=0 timeout 300 python scripts/gpu_mini.py 2>&1 | tail -1
done > gpurun_out/r02l_tier_order.txt 2>&1
cat gpurun_out/r02l_tier_order.txt
