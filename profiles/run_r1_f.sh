#!/usr/bin/env bash
# Round-1 GPU session F: ncu --set full captures of the current whole-step kernels (RK4 fp32, RK2 fp32 4 cells/lane,
# Euler/RK2 fp64 2 cells/lane); each capture only after the same command exited 0 without ncu.
set -u
mkdir -p gpurun_out
cap() {  # name, bench args...
  local name=$1; shift
  local B="python bench.py $* --steps 12 --warmup 3 --no-cpu-baseline --no-e2e"
  $B > gpurun_out/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:step_tma -s 5 -c 1 -f -o gpurun_out/prof_r1f_$name $B > gpurun_out/ncu_$name.log 2>&1
  echo "$name: $?"
}
cap rk4
cap prim --workload prim2048x64
cap baro --workload baro16384_f64
ls -la gpurun_out/*.ncu-rep
