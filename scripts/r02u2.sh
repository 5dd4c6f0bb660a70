#!/bin/bash
# source-level capture of the dominant Jacobian launch (first pass of the <=8-row tier) of one timed step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --profile"
timeout 300 $CMD > gpurun_out/r02u2_plain.json 2> gpurun_out/r02u2_plain.err && \
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:proximity_persistent<\(int\)2>' -s 12 -c 1 -f -o gpurun_out/r02u2_jac_t8 $CMD > gpurun_out/r02u2_ncu.log 2>&1
ls -la gpurun_out/r02u2_*; tail -3 gpurun_out/r02u2_ncu.log
