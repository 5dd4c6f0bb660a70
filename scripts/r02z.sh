#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{ echo "== default"; python scripts/gpu_mini.py | tail -1; echo "== DCOL_SKIP_RESUME=1 (first passes only; timing experiment)"; DCOL_SKIP_RESUME=1 python scripts/gpu_mini.py | tail -1; 
  echo "== DCOL_TIERS=16,32,64,128 (one merged tier)"; DCOL_TIERS=16,32,64,128 python scripts/gpu_mini.py | tail -1;
  echo "== DCOL_TIERS=16,32,64,128 DCOL_SKIP_RESUME=1"; DCOL_TIERS=16,32,64,128 DCOL_SKIP_RESUME=1 python scripts/gpu_mini.py | tail -1; } > gpurun_out/r02z_resume.log 2>&1
timeout 300 ./build_variants/latency_capi > gpurun_out/r02z_latency_capi.log 2>&1
cat gpurun_out/r02z_resume.log gpurun_out/r02z_latency_capi.log
