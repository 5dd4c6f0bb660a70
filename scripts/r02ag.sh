#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "=== $1 [$2]"; env $2 DCOL_B200_LIB=build_variants/libdcol_$1.so timeout 300 python scripts/gpu_mini.py 2>&1 | tail -1; }
{ for v in $VARIANTS; do run $v ""; done; for v in $VARIANTS; do run $v ""; done; } > gpurun_out/${TAG}_variants.log 2>&1
cat gpurun_out/${TAG}_variants.log
