#!/bin/bash
# Development aid: A/B library variants on the GPU box.  usage: gpu_ab.sh <lib.so> [<lib.so> ...]
for L in "$@"; do
  echo "=== $L"
  DCOL_B200_LIB=$L python scripts/gpu_perf.py 2>&1 | awk '{print $1, $2, $5, $6}' | paste -sd' ' | fold -w 200
  for i in 1 2; do DCOL_B200_LIB=$L python scripts/gpu_one.py 1000000 mixed 0 | tail -1; done
  DCOL_B200_LIB=$L python scripts/gpu_one.py 2000000 mixed 2 | tail -1
done
