#!/bin/bash
# round-2 session 3: LDL' factorisation / lean sqrt / register-budget variants (A/B) + GPU tests of the new default build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02u_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02u_tests.log
run() { echo "=== $1 [$2]"; env $2 DCOL_B200_LIB=build_variants/libdcol_$1.so timeout 300 python scripts/gpu_mini.py 2>&1 | tail -1; }
{ run r2final ""; run ldl ""; run sq ""; run both ""; run r224 "DCOL_WARPS=3"; run r200 "DCOL_WARPS=5"; run r2final ""; run both ""; } > gpurun_out/r02u_variants.log 2>&1
perf() { echo "=== $1 [$2]"; env $2 DCOL_B200_LIB=build_variants/libdcol_$1.so timeout 300 python scripts/gpu_perf.py 2>&1 | awk '{print $1, $2, $5, $6}' | paste -sd' ' | fold -w 200; }
{ perf r2final ""; perf both ""; perf r224 "DCOL_WARPS=3"; perf r200 "DCOL_WARPS=5"; } > gpurun_out/r02u_ab.log 2>&1
tail -3 gpurun_out/r02u_tests.log; cat gpurun_out/r02u_variants.log; cat gpurun_out/r02u_ab.log
