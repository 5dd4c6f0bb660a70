#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_group.py tests/test_gpu_hostpipe.py -m gpu -x -q > gpurun_out/r02ac_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ac_tests.log
timeout 600 python scripts/gpu_group.py > gpurun_out/r02ac_group.log 2>&1
tail -4 gpurun_out/r02ac_tests.log; cat gpurun_out/r02ac_group.log
