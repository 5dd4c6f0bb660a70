#!/bin/bash
# round-2 GPU session e: separate pack kernel, mapped error flag, split tiers; bench.py with the new objects; launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_tests.log
run() { echo "=== $1 [$2]"; env $2 DCOL_B200_LIB=build_variants/libdcol_$1.so timeout 300 python scripts/gpu_mini.py 2>&1 | tail -1; }
{ run new2 ""; run new3 ""; } > gpurun_out/r02e_variants.log 2>&1
timeout 300 ./build_variants/latency_capi > gpurun_out/r02e_latency_capi.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02e_bench_ref.json 2>> gpurun_out/r02e_bench.err
timeout 120 python scripts/gpu_one.py 1000000 mixed 0 > gpurun_out/r02e_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02e_launches.csv python scripts/gpu_one.py 1000000 mixed 0 > gpurun_out/r02e_ncu.log 2>&1
tail -4 gpurun_out/r02e_tests.log; cat gpurun_out/r02e_variants.log gpurun_out/r02e_latency_capi.log; tail -5 gpurun_out/r02e_bench.err; cat gpurun_out/r02e_bench.json gpurun_out/r02e_bench_ref.json
