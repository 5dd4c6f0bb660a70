#!/bin/bash
# Round-2 measurement session: tests, the driver's bench command, the reference arm, and the ncu evidence for profiles/
# (launch list of the bench command; --set full captures of one timed step's solver launches; source-level capture of the
# dominant kernel).  Only summaries travel back: the multi-launch .ncu-rep files are reduced to raw CSV on the box.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=/tmp/dcol_prof; mkdir -p $T
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/r02p_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02p_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02p_tests.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02p_bench_ref.json 2>> gpurun_out/r02p_bench.err
timeout 300 ./build_variants/latency_capi > gpurun_out/r02p_latency_capi.log 2>&1
timeout 600 python scripts/gpu_configs.py > gpurun_out/r02p_configs.txt 2>&1
timeout 300 python scripts/gpu_warm.py > gpurun_out/r02p_warm.txt 2>&1
CMD="python bench.py --steps 2 --warmup 3 --profile"
timeout 300 $CMD > gpurun_out/r02p_plain.json 2> gpurun_out/r02p_plain.err && {
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02p_launches.csv $CMD > gpurun_out/r02p_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:proximity_persistent<\(int\)0>' -s 12 -c 1 -f -o gpurun_out/r02p_prox_t8 $CMD >> gpurun_out/r02p_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:proximity_persistent<\(int\)0>' -s 12 -c 4 -f -o $T/prox4 $CMD >> gpurun_out/r02p_ncu.log 2>&1
ncu -i $T/prox4.ncu-rep --page raw --csv > gpurun_out/r02p_prox_step_raw.csv 2>> gpurun_out/r02p_ncu.log
timeout 900 ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:proximity_persistent<\(int\)2>' -s 12 -c 4 -f -o $T/jac4 $CMD >> gpurun_out/r02p_ncu.log 2>&1
ncu -i $T/jac4.ncu-rep --page raw --csv > gpurun_out/r02p_jac_step_raw.csv 2>> gpurun_out/r02p_ncu.log
}
ls -la gpurun_out/r02p_*; tail -3 gpurun_out/r02p_tests.log; tail -3 gpurun_out/r02p_bench.err; cut -c1-600 gpurun_out/r02p_bench.json
