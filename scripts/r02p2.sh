#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02p2_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02p2_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-large --no-cpu-baseline > gpurun_out/r02p2_bench.json 2> gpurun_out/r02p2_bench.err
tail -3 gpurun_out/r02p2_tests.log; python -c "
import json; d=json.loads(open('gpurun_out/r02p2_bench.json').read().strip().split('\n')[-1]); print(json.dumps(d['latency'], indent=0))"
