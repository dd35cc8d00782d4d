#!/usr/bin/env bash
# Round-1 GPU session H: A/B of a kernel variant (lib/ab_*.so) against the current build: parity, then all workloads.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
for w in swe8192_rk4 swe8192_euler prim2048x64 baro16384_f64; do
  timeout 300 bash profiles/ab_workload.sh $w ab_nopure.so libweather_b200.so
done 2>&1 | tee gpurun_out/ab_pure.txt
