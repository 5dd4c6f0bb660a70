#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_operators.py tests/test_gpu_parity.py -m gpu -x -q -k "se3 or se2 or single_pair" > gpurun_out/r02q_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02q_tests.log
timeout 600 python scripts/gpu_configs.py > gpurun_out/r02q_configs.txt 2>&1
tail -12 gpurun_out/r02q_tests.log; cat gpurun_out/r02q_configs.txt
