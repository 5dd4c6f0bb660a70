#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02am_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02am_tests.log
run() { echo "=== $1 [$2]"; env $2 DCOL_B200_LIB=build_variants/libdcol_$1.so timeout 300 python scripts/gpu_mini.py 2>&1 | tail -1; }
{ run fin ""; run sph ""; run fin ""; run sph ""; } > gpurun_out/r02am_variants.log 2>&1
tail -12 gpurun_out/r02am_tests.log; cat gpurun_out/r02am_variants.log
