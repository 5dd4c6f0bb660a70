#!/bin/bash
# host-packed batches of up to 128 pairs (one launch): GPU suite, latency with and without
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02af2_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02af2_tests.log
for k in 1 2; do
echo "=== host-packed up to 128 pairs (default)"; timeout 300 ./build_variants/latency_capi 2>&1 | head -6
echo "=== DCOL_PACKED_MAX=1"; DCOL_PACKED_MAX=1 timeout 300 ./build_variants/latency_capi 2>&1 | head -6
done > gpurun_out/r02af2_latency.txt 2>&1
tail -4 gpurun_out/r02af2_tests.log; cat gpurun_out/r02af2_latency.txt
