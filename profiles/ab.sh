#!/usr/bin/env bash
# A/B of alternative builds of libweather_b200.so: interleaved repetitions of the device-resident bench.
# usage: bash profiles/ab.sh libA.so libB.so [...]   (paths relative to nvidia-jetson-workload_b200/lib)
set -u
for rep in 1 2 3; do for lib in "$@"; do
  WSB_LIBRARY=$PWD/nvidia-jetson-workload_b200/lib/$lib python bench.py --no-cpu-baseline --no-e2e --steps 100 --warmup 10 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$lib rep$rep %.4f ms/step %.2f Gcell/s' % (d['ms_per_step'], d['value']/1e9))"
done; done
