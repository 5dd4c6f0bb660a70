#!/bin/bash
# session 4: GPU tests (with the verbose-trace and staging-pipeline tests), smoke, one verbose table for the docs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02g_smoke.log 2>&1
timeout 120 python - > gpurun_out/r02g_verbose.txt 2>&1 <<'PY'
import numpy as np, dcol_b200 as dc
cone = dc.Cone(2.0, np.deg2rad(22)); cone.r = np.array([0.3, -0.2, 0.1]); cone.q = np.array([0.9, 0.1, -0.3, 0.2])
caps = dc.Capsule(0.3, 1.2); caps.r = np.array([1.5, 0.4, -0.7]); caps.q = np.array([0.7, -0.5, 0.2, 0.4])
print(dc.proximity(cone, caps, verbose=True))
PY
tail -5 gpurun_out/r02g_tests.log; cat gpurun_out/r02g_smoke.log | tail -3; cat gpurun_out/r02g_verbose.txt
