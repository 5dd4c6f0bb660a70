#!/bin/bash
# all visible GPUs: sharded scene (group + torchrun ranks), then both bench arms under torchrun
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 600 python scripts/gpu_scene_multi.py > gpurun_out/r02n_scene_group_${NG}gpu.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29611 scripts/gpu_scene_multi.py > gpurun_out/r02n_scene_torchrun_${NG}gpu.txt 2>&1
bash scripts/r02m.sh > /dev/null 2>&1
cp gpurun_out/r02m_bench_${NG}gpu.json gpurun_out/r02n_bench_${NG}gpu.json; cp gpurun_out/r02m_ref_${NG}gpu.json gpurun_out/r02n_ref_${NG}gpu.json
cat gpurun_out/r02n_scene_group_${NG}gpu.txt; grep "^torchrun" gpurun_out/r02n_scene_torchrun_${NG}gpu.txt; cut -c1-200 gpurun_out/r02n_bench_${NG}gpu.json; tail -2 gpurun_out/r02m_bench_${NG}gpu.err
