#!/bin/bash
# Round-2 measurement session: tests, the driver's bench command, the reference arm, and the ncu evidence for profiles/
# (launch list of the bench command; --set full captures of one timed step's solver launches; source-level capture of the
# dominant kernel).  Only summaries travel back: the multi-launch .ncu-rep files are reduced to raw CSV on the box.
cd "$(dirname "$0")/.."
TAG=${1:-r02p}
mkdir -p gpurun_out
T=/tmp/dcol_prof; mkdir -p $T
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
timeout 300 ./build_variants/latency_capi > gpurun_out/${TAG}_latency_capi.log 2>&1
timeout 600 python scripts/gpu_configs.py > gpurun_out/${TAG}_configs.txt 2>&1
timeout 300 python scripts/gpu_warm.py > gpurun_out/${TAG}_warm.txt 2>&1
timeout 600 python scripts/gpu_hostpipe.py > gpurun_out/${TAG}_hostpipe.txt 2>&1
CMD="python bench.py --steps 2 --warmup 3 --profile"
timeout 300 $CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err && {
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:proximity_persistent<\(int\)0>' -s 12 -c 1 -f -o gpurun_out/${TAG}_prox_t8 $CMD >> gpurun_out/${TAG}_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:proximity_persistent<\(int\)0>' -s 12 -c 4 -f -o $T/prox4 $CMD >> gpurun_out/${TAG}_ncu.log 2>&1
ncu -i $T/prox4.ncu-rep --page raw --csv > gpurun_out/${TAG}_prox_step_raw.csv 2>> gpurun_out/${TAG}_ncu.log
timeout 900 ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:proximity_persistent<\(int\)2>' -s 12 -c 4 -f -o $T/jac4 $CMD >> gpurun_out/${TAG}_ncu.log 2>&1
ncu -i $T/jac4.ncu-rep --page raw --csv > gpurun_out/${TAG}_jac_step_raw.csv 2>> gpurun_out/${TAG}_ncu.log
}
ls -la gpurun_out/${TAG}_*; tail -3 gpurun_out/${TAG}_tests.log; tail -3 gpurun_out/${TAG}_bench.err; cut -c1-600 gpurun_out/${TAG}_bench.json
