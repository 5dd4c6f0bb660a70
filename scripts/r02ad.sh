#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "=== $1 [$2]"; env $2 DCOL_B200_LIB=build_variants/libdcol_$1.so timeout 300 python scripts/gpu_mini.py 2>&1 | tail -1; }
{ run cur ""; run lsq ""; run lsq2 ""; run cur ""; run lsq ""; run lsq2 ""; } > gpurun_out/r02ad_variants.log 2>&1
perf() { echo "=== $1 [$2]"; env $2 DCOL_B200_LIB=build_variants/libdcol_$1.so timeout 300 python scripts/gpu_perf.py 2>&1 | awk '{print $1, $2, $5, $6}' | paste -sd' ' | fold -w 200; }
{ perf lsq ""; perf lsq2 ""; } > gpurun_out/r02ad_ab.log 2>&1
cat gpurun_out/r02ad_variants.log gpurun_out/r02ad_ab.log
