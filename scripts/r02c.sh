#!/bin/bash
# round-2 GPU session c: packed-record data path + small-batch path; 168-register variants; parity table; latency
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_tests.log
timeout 900 bash scripts/gpu_ab.sh build_variants/libdcol_wp.so build_variants/libdcol_wp168.so build_variants/libdcol_new.so build_variants/libdcol_new168.so > gpurun_out/r02c_ab.log 2>&1
timeout 600 python scripts/gpu_latency.py > gpurun_out/r02c_latency.log 2>&1
timeout 1500 python scripts/parity_table.py 100000 > gpurun_out/r02c_parity.log 2>&1
export DCOL_B200_LIB=build_variants/libdcol_new.so
timeout 120 python scripts/gpu_one.py 1000000 mixed 0 > gpurun_out/r02c_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:proximity_persistent -s 2 -c 1 -o gpurun_out/r02c_prof -f python scripts/gpu_one.py 1000000 mixed 0 > gpurun_out/r02c_ncu.log 2>&1
ls -la gpurun_out/ | tail -12
tail -5 gpurun_out/r02c_tests.log; cat gpurun_out/r02c_ab.log | fold -w 200; cat gpurun_out/r02c_latency.log; cat gpurun_out/r02c_parity.log
