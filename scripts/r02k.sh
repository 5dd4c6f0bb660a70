#!/bin/bash
# session 4 final: the driver's bench command (with the large_batch object), the default command, the reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T0=$(date +%s.%N)
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
T1=$(date +%s.%N)
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02k_bench_ref.json 2>> gpurun_out/r02k_bench.err
T2=$(date +%s.%N)
timeout 900 python bench.py > gpurun_out/r02k_bench_default.json 2> gpurun_out/r02k_bench_default.err
T3=$(date +%s.%N)
echo "wall seconds: bench --steps 20 --warmup 3: $(echo "$T1 - $T0" | bc), --impl reference: $(echo "$T2 - $T1" | bc), bench (defaults): $(echo "$T3 - $T2" | bc)" > gpurun_out/r02k_wall.txt
cat gpurun_out/r02k_wall.txt; tail -3 gpurun_out/r02k_bench.err; cut -c1-300 gpurun_out/r02k_bench.json
