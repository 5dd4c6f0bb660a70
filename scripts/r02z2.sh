#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "small or single or edge or config1 or latency or golden" > gpurun_out/r02z2_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02z2_tests.log
for k in 1 2 3; do
echo "=== 128-thread preparation for n <= 128 (default)"; timeout 300 ./build_variants/latency_capi 2>&1 | head -6
echo "=== DCOL_PREP_1024=1"; DCOL_PREP_1024=1 timeout 300 ./build_variants/latency_capi 2>&1 | head -6
done > gpurun_out/r02z2_latency.txt 2>&1
tail -3 gpurun_out/r02z2_tests.log; cat gpurun_out/r02z2_latency.txt
