#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "=== $1 [$2]"; env $2 DCOL_B200_LIB=build_variants/libdcol_$1.so timeout 300 python scripts/gpu_mini.py 2>&1 | tail -1; }
{ run lss ""; run jdk ""; run etaf ""; run lss ""; run jdk ""; run etaf ""; } > gpurun_out/r02af_variants.log 2>&1
DCOL_B200_LIB=build_variants/libdcol_etaf.so timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02af_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02af_tests.log
cat gpurun_out/r02af_variants.log; tail -12 gpurun_out/r02af_tests.log
