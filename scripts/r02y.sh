#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python scripts/gpu_hostpipe_sizes.py > gpurun_out/r02y_sizes.log 2>&1
timeout 300 ./build_variants/latency_capi 2>&1 | grep 100000 > gpurun_out/r02y_latency.log
timeout 600 python -m pytest tests/test_gpu_hostpipe.py -m gpu -x -q > gpurun_out/r02y_tests.log 2>&1
cat gpurun_out/r02y_sizes.log gpurun_out/r02y_latency.log; tail -3 gpurun_out/r02y_tests.log
