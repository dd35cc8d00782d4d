#!/usr/bin/env bash
# rows-per-chunk sweep for the HBM-bound members of the whole-step family
set -u
mkdir -p gpurun_out
for w in swe8192_euler baro16384_f64 prim2048x64; do for rpc in 32 64 128 256 512; do
  WSB_FUSED_ROWS_PER_CHUNK=$rpc python bench.py --workload $w --steps 20 --warmup 3 --no-e2e --no-cpu-baseline | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$w rpc=$rpc %.4f ms/step %.1f Gcell/s %.0f GB/s frac=%.3f' % (d['ms_per_step'], d['value']/1e9, r['achieved'], r['frac']))"
done; done 2>&1 | tee gpurun_out/sweep_rpc_others.log
