#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02x_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02x_tests.log
timeout 300 ./build_variants/latency_capi > gpurun_out/r02x_latency_capi.log 2>&1
DCOL_SMALL_MERGED=0 timeout 300 ./build_variants/latency_capi > gpurun_out/r02x_latency_capi_unmerged.log 2>&1
tail -15 gpurun_out/r02x_tests.log; cat gpurun_out/r02x_latency_capi.log; echo unmerged; cat gpurun_out/r02x_latency_capi_unmerged.log
