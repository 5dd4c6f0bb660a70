#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02ab2_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ab2_tests.log
for k in 1 2 3; do timeout 300 ./build_variants/latency_capi 2>&1 | head -8; echo; done > gpurun_out/r02ab2_latency.txt 2>&1
timeout 300 ./build_variants/latency_capi > gpurun_out/r02ab2_latency_full.txt 2>&1
tail -3 gpurun_out/r02ab2_tests.log; cat gpurun_out/r02ab2_latency.txt
