#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_tests.log
timeout 300 python scripts/gpu_mini.py > gpurun_out/r02h_mini.txt 2>&1
timeout 300 ./build_variants/latency_capi > gpurun_out/r02h_latency_capi.log 2>&1
timeout 120 python scripts/gpu_one.py 1000000 mixed 0 > gpurun_out/r02h_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02h_launches.csv python scripts/gpu_one.py 1000000 mixed 0 > gpurun_out/r02h_ncu.log 2>&1
tail -3 gpurun_out/r02h_tests.log; cat gpurun_out/r02h_mini.txt gpurun_out/r02h_latency_capi.log
