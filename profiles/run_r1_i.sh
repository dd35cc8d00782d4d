#!/usr/bin/env bash
# Round-1 GPU session I: fill-skip A/B and rows-per-chunk sweep of the RK4 kernel after it became fp32-pipe bound.
set -u
mkdir -p gpurun_out
WSB_LIBRARY=$PWD/nvidia-jetson-workload_b200/lib/ab_fs1.so timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -2
bash profiles/ab.sh ab_fs0.so ab_fs1.so 2>&1 | tee gpurun_out/ab_fillskip.txt
for lib in ab_fs0.so ab_fs1.so; do for rpc in 48 64 96 128 192; do
  WSB_FUSED_ROWS_PER_CHUNK=$rpc WSB_LIBRARY=$PWD/nvidia-jetson-workload_b200/lib/$lib python bench.py --no-cpu-baseline --no-e2e --steps 100 --warmup 10 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$lib rpc=$rpc %.4f ms/step %.2f Gcell/s' % (d['ms_per_step'], d['value']/1e9))"
done; done 2>&1 | tee gpurun_out/sweep_rpc_r1i.txt
