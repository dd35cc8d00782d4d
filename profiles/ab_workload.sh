#!/usr/bin/env bash
# A/B of alternative builds on one workload: bash profiles/ab_workload.sh <workload> libA.so libB.so ...
set -u
w=$1; shift
for rep in 1 2 3; do for lib in "$@"; do
  WSB_LIBRARY=$PWD/nvidia-jetson-workload_b200/lib/$lib python bench.py --workload $w --no-cpu-baseline --no-e2e --steps 50 --warmup 10 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$w $lib rep$rep %.4f ms/step %.2f Gcell/s frac %.3f clocks %s' % (d['ms_per_step'], d['value']/1e9, d['roofline']['frac'], d['clocks']))"
done; done
