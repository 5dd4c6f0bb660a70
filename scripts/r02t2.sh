#!/bin/bash
# FP64 peak probe, both forms, plain and under ncu; then the driver's bench command with the pipe peak as the denominator
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/dfma.py <<'PY'
import sys; sys.path.insert(0, '.')
import dcol_b200 as dc
ctx = dc.Context(0)
print("x = fma(x, a, x), 0.3 s:", ctx.measure_fp64_tflops(0.3), "TFLOP/s")
print("x = fma(x, a, b), 0.3 s:", ctx.measure_fp64_tflops(0.3, three_source=True), "TFLOP/s")
ctx.close()
PY
timeout 120 python /tmp/dfma.py > gpurun_out/r02t2_dfma_plain.txt 2>&1 && \
timeout 600 ncu --clock-control none -k regex:k_dfma -c 400 --metrics sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second --csv python /tmp/dfma.py > gpurun_out/r02t2_dfma_ncu.csv 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02t2_bench.json 2> gpurun_out/r02t2_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02t2_bench_ref.json 2>> gpurun_out/r02t2_bench.err
cat gpurun_out/r02t2_dfma_plain.txt; tail -2 gpurun_out/r02t2_bench.err
python - <<'PY'
import csv, json, collections
rows = [r for r in csv.reader(open('gpurun_out/r02t2_dfma_ncu.csv', errors='replace')) if len(r) > 10]
hdr = rows[0]; agg = collections.defaultdict(list)
for r in rows[1:]:
    d = dict(zip(hdr, r))
    try: agg[(d['Kernel Name'][:40], d['Metric Name'])].append(float(d['Metric Value'].replace(',', '')))
    except Exception: pass
for k, v in sorted(agg.items()):
    v = sorted(v); print(k, 'n=%d median=%.2f' % (len(v), v[len(v) // 2]))
d = json.loads(open('gpurun_out/r02t2_bench.json').read().strip().split('\n')[-1])
r = d['roofline']; print({k: r[k] for k in ('achieved', 'peak', 'frac', 'frac_of_nominal_37.2', 'peak_three_source_dfma', 'frac_of_three_source_peak')}, d['value'], d['large_batch']['roofline_frac'], d['jacobian']['roofline']['frac'])
PY
