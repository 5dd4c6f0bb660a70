#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_tests.log
timeout 300 python scripts/gpu_warm.py > gpurun_out/r02g_warm.txt 2>&1
timeout 300 python scripts/gpu_mini.py > gpurun_out/r02g_mini.txt 2>&1
tail -15 gpurun_out/r02g_tests.log; cat gpurun_out/r02g_warm.txt gpurun_out/r02g_mini.txt
