#!/bin/bash
# session 4: maximum-size call, compute-sanitizer attempt (one GPU)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python scripts/gpu_big.py > gpurun_out/r02i_big.txt 2>&1; echo "rc=$?" >> gpurun_out/r02i_big.txt
timeout 120 python scripts/sanitize_small.py > gpurun_out/r02i_sanitize_plain.txt 2>&1; echo "rc=$?" >> gpurun_out/r02i_sanitize_plain.txt
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/r02i_memcheck.txt 2>&1; echo "rc=$?" >> gpurun_out/r02i_memcheck.txt
timeout 420 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/r02i_racecheck.txt 2>&1; echo "rc=$?" >> gpurun_out/r02i_racecheck.txt
tail -6 gpurun_out/r02i_big.txt; tail -3 gpurun_out/r02i_sanitize_plain.txt; tail -8 gpurun_out/r02i_memcheck.txt; tail -8 gpurun_out/r02i_racecheck.txt
