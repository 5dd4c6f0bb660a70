#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02w_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02w_tests.log
run() { echo "=== $1 [$2]"; env $2 DCOL_B200_LIB=build_variants/libdcol_$1.so timeout 300 python scripts/gpu_mini.py 2>&1 | tail -1; }
{ run r2final ""; run both7 ""; run r2final ""; run both7 ""; } > gpurun_out/r02w_variants.log 2>&1
perf() { echo "=== $1 [$2]"; env $2 DCOL_B200_LIB=build_variants/libdcol_$1.so timeout 300 python scripts/gpu_perf.py 2>&1 | awk '{print $1, $2, $5, $6}' | paste -sd' ' | fold -w 200; }
{ perf both7 ""; } > gpurun_out/r02w_ab.log 2>&1
tail -15 gpurun_out/r02w_tests.log; cat gpurun_out/r02w_variants.log gpurun_out/r02w_ab.log
