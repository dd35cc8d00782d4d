#!/usr/bin/env bash
# Round-1 GPU session G: after the producer / pointer changes: other workloads + full capture of the RK4 kernel.
set -u
mkdir -p gpurun_out
for w in swe8192_euler prim2048x64 baro16384_f64; do
  python bench.py --workload $w --no-cpu-baseline --no-e2e --steps 100 --warmup 10 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$w %.4f ms/step %.2f Gcell/s frac %.3f' % (d['ms_per_step'], d['value']/1e9, d['roofline']['frac']))"
done 2>&1 | tee gpurun_out/others_g.txt
B="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e"
$B > gpurun_out/plain_g.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_tma -s 5 -c 1 -f -o gpurun_out/prof_r1g_rk4 $B > gpurun_out/ncu_g.log 2>&1
echo "ncu: $?"
