#!/bin/bash
# final check of the round on the committed build: GPU suite, smoke, both bench arms
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02final_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02final_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02final_smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02final_bench.json 2> gpurun_out/r02final_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02final_bench_ref.json 2>> gpurun_out/r02final_bench.err
tail -3 gpurun_out/r02final_tests.log; tail -1 gpurun_out/r02final_smoke.log; tail -2 gpurun_out/r02final_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r02final_bench.json').read().strip().split('\n')[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['jacobian']['value'], d['large_batch']['value'], d['latency']['single_pair_proximity'], d['clocks']); r=json.loads(open('gpurun_out/r02final_bench_ref.json').read().strip().split('\n')[-1]); print(r['value'], r['cpu_baseline']['cores'])"
