#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02ae_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ae_tests.log
run() { echo "=== $1 [$2]"; env $2 DCOL_B200_LIB=build_variants/libdcol_$1.so timeout 300 python scripts/gpu_mini.py 2>&1 | tail -1; }
{ run lsq ""; run lss ""; run lsq ""; run lss ""; } > gpurun_out/r02ae_variants.log 2>&1
perf() { echo "=== $1 [$2]"; env $2 DCOL_B200_LIB=build_variants/libdcol_$1.so timeout 300 python scripts/gpu_perf.py 2>&1 | awk '{print $1, $2, $5, $6}' | paste -sd' ' | fold -w 200; }
{ perf lss ""; } > gpurun_out/r02ae_ab.log 2>&1
timeout 600 python scripts/gpu_configs.py 2>&1 | grep "config 4" > gpurun_out/r02ae_configs.log
tail -12 gpurun_out/r02ae_tests.log; cat gpurun_out/r02ae_variants.log gpurun_out/r02ae_ab.log gpurun_out/r02ae_configs.log
