#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
./build_variants/dfma_probe > gpurun_out/r02s2_probe_plain.txt 2>&1
timeout 600 ncu --clock-control none --metrics sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active --csv ./build_variants/dfma_probe > gpurun_out/r02s2_probe_ncu.csv 2>&1
cat gpurun_out/r02s2_probe_plain.txt; grep -v "^==" gpurun_out/r02s2_probe_ncu.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=None
for r in rows:
    if 'Metric Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        if d['Metric Name'].startswith('sm__pipe_fp64') or d['Metric Name'].startswith('gpu__time'): print(d['ID'], d['Kernel Name'][:40], d['Grid Size'], d['Block Size'], d['Metric Name'][:30], d['Metric Value'])
"
