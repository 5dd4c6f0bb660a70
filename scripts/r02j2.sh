#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python scripts/gpu_sizes.py > gpurun_out/r02j_sizes.txt 2>&1
DCOL_SKIP_RESUME=1 timeout 600 python scripts/gpu_sizes.py > gpurun_out/r02j_sizes_noresume.txt 2>&1
cat gpurun_out/r02j_sizes.txt; echo ---- no resume; cat gpurun_out/r02j_sizes_noresume.txt
