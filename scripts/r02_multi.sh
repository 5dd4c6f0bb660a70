#!/bin/bash
# usage: r02_multi.sh <n_gpus> <tag>   (run under gpurun --gpus <n_gpus>)
cd "$(dirname "$0")/.."
N=$1; TAG=$2
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
tail -2 gpurun_out/${TAG}_bench_${N}gpu.err; cut -c1-900 gpurun_out/${TAG}_bench_${N}gpu.json
