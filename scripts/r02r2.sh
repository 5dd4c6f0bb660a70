#!/bin/bash
# FP64 peak probe under ncu: is the DFMA pipe saturated by k_dfma (roofline denominator)?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/dfma.py <<'PY'
import sys; sys.path.insert(0, '.')
import dcol_b200 as dc
ctx = dc.Context(0)
print("measured FP64 peak (0.05 s):", ctx._lib.dcol_measure_fp64_tflops_for(ctx._h, 0.05), "TFLOP/s")
ctx.close()
PY
timeout 120 python /tmp/dfma.py > gpurun_out/r02r2_dfma_plain.txt 2>&1 && \
timeout 600 ncu --clock-control none -k regex:k_dfma -s 2 -c 2 --metrics sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active python /tmp/dfma.py > gpurun_out/r02r2_dfma_ncu.txt 2>&1
cat gpurun_out/r02r2_dfma_plain.txt; grep -v "^==PROF==" gpurun_out/r02r2_dfma_ncu.txt | tail -30
