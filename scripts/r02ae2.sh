#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02ae2_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ae2_tests.log
timeout 300 ./build_variants/latency_capi > gpurun_out/r02ae2_latency_capi.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02ae2_bench.json 2> gpurun_out/r02ae2_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02ae2_bench_ref.json 2>> gpurun_out/r02ae2_bench.err
tail -3 gpurun_out/r02ae2_tests.log; cat gpurun_out/r02ae2_latency_capi.log; python -c "
import json; d=json.loads(open('gpurun_out/r02ae2_bench.json').read().strip().split('\n')[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['jacobian']['value']); print(json.dumps(d['latency'], indent=0))"
