#!/bin/bash
# final pass of round 2, session 3: parity table, then the measurement session (scripts/r02_profile.sh)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python scripts/parity_table.py 100000 > gpurun_out/r02f_parity.log 2>&1
bash scripts/r02_profile.sh r02f
tail -30 gpurun_out/r02f_parity.log
