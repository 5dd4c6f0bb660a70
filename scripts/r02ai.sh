#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --profile"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:proximity_persistent<\(int\)2>' -s 12 -c 1 -f -o gpurun_out/r02ai_jac_t8 $CMD > gpurun_out/r02ai_ncu.log 2>&1
ls -la gpurun_out/r02ai_*; tail -3 gpurun_out/r02ai_ncu.log
