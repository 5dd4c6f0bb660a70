#!/bin/bash
# final small-call latency (zero-copy staging up to 1024 pairs): C ABI tool, Python wrapper (bench latency object), small-path tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02y2_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02y2_tests.log
timeout 300 ./build_variants/latency_capi > gpurun_out/r02y2_latency_capi.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-large --no-cpu-baseline --no-jacobian > gpurun_out/r02y2_bench.json 2> gpurun_out/r02y2_bench.err
tail -3 gpurun_out/r02y2_tests.log; cat gpurun_out/r02y2_latency_capi.log; python -c "
import json; d=json.loads(open('gpurun_out/r02y2_bench.json').read().strip().split('\n')[-1]); print(json.dumps(d['latency'], indent=0))"
