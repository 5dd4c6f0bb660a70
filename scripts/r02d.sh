#!/bin/bash
# round-2 GPU session d: micro-variants (tiers, row unrolling, lockstep, hand-over threshold, warps), latency, tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_tests.log
run() { echo "=== $1 [$2]"; env $2 DCOL_B200_LIB=build_variants/libdcol_$1.so timeout 300 python scripts/gpu_mini.py 2>&1 | tail -1; }
{
run wp ""
run new2 ""
run new2 "DCOL_TIERS=8,16,32,64,128"
run new2 "DCOL_SPILL=4"
run new2 "DCOL_SPILL=6"
run new2 "DCOL_SPILL=12"
run new2 "DCOL_SPILL=16"
run new2 "DCOL_SPILL=0"
run new2 "DCOL_WARPS=8"
run new2 "DCOL_WARPS=2"
run new2 "DCOL_LEVELS=2"
run unroll2 ""
run nolock ""
} > gpurun_out/r02d_variants.log 2>&1
timeout 300 ./build_variants/latency_capi > gpurun_out/r02d_latency_capi.log 2>&1
timeout 600 python scripts/gpu_latency.py > gpurun_out/r02d_latency_py.log 2>&1
tail -4 gpurun_out/r02d_tests.log; cat gpurun_out/r02d_variants.log gpurun_out/r02d_latency_capi.log gpurun_out/r02d_latency_py.log
