#!/usr/bin/env bash
# r1: host-buffer step: kernel stores straight to pinned host memory (experiment) vs D2H copies.
set -u
mkdir -p gpurun_out
for zc in 1 0; do for rpc in 16 64; do for n in 8 16 32 64; do
  if [ $zc = 1 ]; then export WSB_HOST_ZEROCOPY=1; else unset WSB_HOST_ZEROCOPY; fi
  WSB_HOST_RPC=$rpc WSB_HOST_SLABS=$n timeout 120 python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('zc=$zc rpc=$rpc slabs=$n e2e %.3f ms/step %.3f Gcell/s' % (d['e2e']['ms_per_step'], d['e2e']['value']/1e9))"
done; done; done 2>&1 | tee gpurun_out/e2e_sweep3.txt
