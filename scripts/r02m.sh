#!/bin/bash
# bench under torchrun on every visible GPU (the driver's N>1 launch), both arms
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $NG --steps 20 --warmup 3 > gpurun_out/r02m_bench_${NG}gpu.json 2> gpurun_out/r02m_bench_${NG}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29712 bench.py --impl reference --gpus $NG --steps 3 --warmup 1 > gpurun_out/r02m_ref_${NG}gpu.json 2>> gpurun_out/r02m_bench_${NG}gpu.err
tail -3 gpurun_out/r02m_bench_${NG}gpu.err; cut -c1-400 gpurun_out/r02m_bench_${NG}gpu.json; cut -c1-200 gpurun_out/r02m_ref_${NG}gpu.json
