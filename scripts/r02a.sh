#!/bin/bash
# round-2 first GPU session: tests, A/B of library variants, ncu of the current solver (source-level)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_tests.log
timeout 600 bash scripts/gpu_ab.sh build_variants/libdcol_base.so build_variants/libdcol_inl0.so > gpurun_out/r02a_ab.log 2>&1
export DCOL_B200_LIB=build_variants/libdcol_base.so
timeout 120 python scripts/gpu_one.py 1000000 mixed 0 > gpurun_out/r02a_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:proximity_persistent -s 4 -c 4 -o gpurun_out/r02a_prof -f python scripts/gpu_one.py 1000000 mixed 0 > gpurun_out/r02a_ncu.log 2>&1
ls -la gpurun_out/r02a_* 
tail -3 gpurun_out/r02a_tests.log; cat gpurun_out/r02a_ab.log | fold -w 200
