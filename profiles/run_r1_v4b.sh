#!/usr/bin/env bash
# r1: RK4 four-cells-per-lane A/B + parity under the override.
set -u
mkdir -p gpurun_out
WSB_CELLS_PER_LANE=4 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -5
for rep in 1 2 3; do for v in 4 2; do for w in swe8192_rk4; do
  WSB_CELLS_PER_LANE=$v python bench.py --workload $w --no-cpu-baseline --no-e2e --steps 100 --warmup 10 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$w cells=$v rep$rep %.4f ms/step %.2f Gcell/s frac %.3f' % (d['ms_per_step'], d['value']/1e9, d['roofline']['frac']))"
done; done; done 2>&1 | tee gpurun_out/v4b_ab.txt
