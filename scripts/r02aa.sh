#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
perf() { echo "=== $1 [$2]"; env $2 DCOL_B200_LIB=build_variants/libdcol_$1.so timeout 300 python scripts/gpu_perf.py 2>&1 | awk '{print $1, $2, $5, $6}' | paste -sd' ' | fold -w 200; }
{ perf cur ""; perf r168 ""; perf r192 ""; perf r168 "DCOL_TIERS=8,12,16,32,64,128"; } > gpurun_out/r02aa_ab.log 2>&1
cat gpurun_out/r02aa_ab.log
