#!/bin/bash
# session 4: sharded all-pairs scene -- tests, then config 4 through a device group and through torchrun ranks
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02h_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_operators.py -m gpu -x -q > gpurun_out/r02h_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_tests.log
timeout 600 python scripts/gpu_scene_multi.py > gpurun_out/r02h_scene_group.txt 2>&1
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29611 scripts/gpu_scene_multi.py > gpurun_out/r02h_scene_torchrun.txt 2>&1
fi
tail -5 gpurun_out/r02h_tests.log; cat gpurun_out/r02h_scene_group.txt; grep -v "^\[W\|^W1\|^\*\*\*" gpurun_out/r02h_scene_torchrun.txt | tail -8
