#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err
CMD="python bench.py --steps 2 --warmup 3 --profile --no-jacobian"
timeout 300 $CMD > gpurun_out/r02s_plain.json 2> gpurun_out/r02s_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k k_dfma -c 3 --csv --log-file gpurun_out/r02s_dfma.csv $CMD > gpurun_out/r02s_ncu.log 2>&1
cat gpurun_out/r02s_dfma.csv | tail -12; cut -c1-300 gpurun_out/r02s_bench.json
