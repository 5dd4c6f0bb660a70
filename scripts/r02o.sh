#!/bin/bash
# flakiness check: the GPU suite five times in a row on one box (hand-over atomics, stream / event choreography, copy threads)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/r02o_repeat.txt
for k in 1 2 3 4 5; do
  timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -2 >> gpurun_out/r02o_repeat.txt
done
cat gpurun_out/r02o_repeat.txt
