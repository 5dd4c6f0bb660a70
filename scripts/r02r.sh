#!/bin/bash
# round-2 hygiene captures: k_dfma under ncu (is the roofline denominator a full FP64 pipe?), DRAM bytes of every kernel of a step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --profile --no-jacobian"
timeout 300 $CMD > gpurun_out/r02r_plain.json 2> gpurun_out/r02r_plain.err && {
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 200 --csv --log-file gpurun_out/r02r_step_metrics.csv $CMD > gpurun_out/r02r_ncu.log 2>&1
}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
