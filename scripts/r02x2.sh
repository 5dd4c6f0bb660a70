#!/bin/bash
# zero-copy staging of tiny host calls: GPU suite, then the C-ABI latency tool with and without it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02x2_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02x2_tests.log
for k in 1 2; do
echo "=== zero-copy for n <= 64 (default)"; timeout 300 ./build_variants/latency_capi 2>&1 | head -8
echo "=== DCOL_ZEROCOPY_MAX=0"; DCOL_ZEROCOPY_MAX=0 timeout 300 ./build_variants/latency_capi 2>&1 | head -8
echo "=== DCOL_ZEROCOPY_MAX=1000"; DCOL_ZEROCOPY_MAX=1000 timeout 300 ./build_variants/latency_capi 2>&1 | head -8
done > gpurun_out/r02x2_latency.txt 2>&1
tail -4 gpurun_out/r02x2_tests.log; cat gpurun_out/r02x2_latency.txt
