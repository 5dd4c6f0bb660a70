#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_tests.log
timeout 600 python scripts/gpu_configs.py > gpurun_out/r02f_configs.txt 2>&1
tail -15 gpurun_out/r02f_tests.log; cat gpurun_out/r02f_configs.txt
