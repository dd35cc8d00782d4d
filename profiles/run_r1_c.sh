#!/usr/bin/env bash
# Round-1 GPU session C: the other BASELINE configurations (device-resident throughput only).
set -u
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e"
for w in swe8192_euler baro16384_f64 prim2048x64 swe32768_rk4; do
  for k in auto stage_direct; do
    echo "== $w kernel=$k"
    $B --workload $w --kernel $k --steps 20 --warmup 3 2> gpurun_out/err_${w}_${k}.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']
    print('%.4f ms/step  %.2f Gcell/s  kernel=%s  achieved=%.0f GB/s frac=%.3f (stage-equivalent %.0f GB/s) clocks=%s' % (d['ms_per_step'], d['value']/1e9, r['kernel'], r['achieved'], r['frac'], r['stage_per_pass_equivalent_gbs'], d['clocks']))"
  done
done 2>&1 | tee gpurun_out/other_workloads.log
