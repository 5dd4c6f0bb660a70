#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02aa2_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02aa2_tests.log
for k in 1 2 3; do
echo "=== early (programmatic) launch of the solver behind the preparation kernel (default)"; timeout 300 ./build_variants/latency_capi 2>&1 | head -8
echo "=== DCOL_PDL=0"; DCOL_PDL=0 timeout 300 ./build_variants/latency_capi 2>&1 | head -8
done > gpurun_out/r02aa2_latency.txt 2>&1
tail -3 gpurun_out/r02aa2_tests.log; cat gpurun_out/r02aa2_latency.txt
