#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02ab_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ab_tests.log
tail -15 gpurun_out/r02ab_tests.log
