#!/usr/bin/env bash
# Round-2 GPU session L (1 GPU): A/B of the early ring refill (y row 3q-4 carried in registers so that the group of
# triple q-2 is re-armed at the top of a steady triple); lib/libweather_b200_norefill.so = same tree without it.
set -u
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e --no-other-configs"
line() { python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$1 %.4f ms/step %.2f Gcell/s frac %.3f clk %s' % (d['ms_per_step'], d['value']/1e9, d['roofline']['frac'], d['clocks']['sm_mhz']))"; }
{
for rep in 1 2 3; do
  WSB_LIBRARY=$PWD/nvidia-jetson-workload_b200/lib/libweather_b200_norefill.so $B --steps 100 --warmup 10 | line "rk4 strict refill after iteration 0 rep$rep"
  $B --steps 100 --warmup 10 | line "rk4 strict early refill            rep$rep"
done
WSB_LIBRARY=$PWD/nvidia-jetson-workload_b200/lib/libweather_b200_norefill.so $B --arith folded --steps 100 --warmup 10 | line "rk4 folded refill after iteration 0"
$B --arith folded --steps 100 --warmup 10 | line "rk4 folded early refill"
} 2>&1 | tee gpurun_out/r2l_ab.txt
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/r2l_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2l_pytest.log
tail -4 gpurun_out/r2l_pytest.log
