#!/bin/bash
# single pair on host pointers: one launch on a host-written record (proximity_single) against the two-launch small path
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02ad2_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ad2_tests.log
for k in 1 2 3; do
echo "=== one launch (default)"; timeout 300 ./build_variants/latency_capi 2>&1 | head -4
echo "=== DCOL_SINGLE=0"; DCOL_SINGLE=0 timeout 300 ./build_variants/latency_capi 2>&1 | head -4
done > gpurun_out/r02ad2_latency.txt 2>&1
tail -3 gpurun_out/r02ad2_tests.log; cat gpurun_out/r02ad2_latency.txt
