#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
perf() { echo "=== $1 [$2]"; env $2 DCOL_B200_LIB=build_variants/libdcol_$1.so timeout 300 python scripts/gpu_perf.py 2>&1 | awk '{print $1, $2, $5, $6}' | paste -sd' ' | fold -w 200; }
{ perf ldl ""; perf sq ""; } > gpurun_out/r02v_ab.log 2>&1
timeout 900 python scripts/gpu_hostpipe.py > gpurun_out/r02v_hostpipe.log 2>&1
cat gpurun_out/r02v_ab.log gpurun_out/r02v_hostpipe.log
