#!/bin/bash
# round-2 GPU session b: whole-program build vs inline-all vs round-1 device-linked build; ncu of the new build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_tests.log
timeout 900 bash scripts/gpu_ab.sh build_variants/libdcol_base.so build_variants/libdcol_inl0.so build_variants/libdcol_wp.so > gpurun_out/r02b_ab.log 2>&1
export DCOL_B200_LIB=build_variants/libdcol_wp.so
timeout 120 python scripts/gpu_one.py 1000000 mixed 0 > gpurun_out/r02b_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:proximity_persistent -s 4 -c 1 -o gpurun_out/r02b_prof_t0 -f python scripts/gpu_one.py 1000000 mixed 0 > gpurun_out/r02b_ncu.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:proximity_persistent -s 6 -c 1 -o gpurun_out/r02b_prof_t1 -f python scripts/gpu_one.py 1000000 mixed 0 >> gpurun_out/r02b_ncu.log 2>&1
ls -la gpurun_out/
tail -3 gpurun_out/r02b_tests.log; cat gpurun_out/r02b_ab.log | fold -w 200
