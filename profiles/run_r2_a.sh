#!/usr/bin/env bash
# Round-2 GPU session A: full parity suite on the packed-FFMA2(-1) arithmetic, A/B against the round-1 library
# (lib/libweather_b200_r1.so, built from commit 5dc98d3) and against the folded opt-in, other workloads, one full
# ncu capture of the RK4 kernel + launch list.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
B="python bench.py --no-cpu-baseline --no-e2e"
line() { python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$1 %.4f ms/step %.2f Gcell/s frac %.3f clk %s' % (d['ms_per_step'], d['value']/1e9, d['roofline']['frac'], d['clocks']['sm_mhz']))"; }
{
for rep in 1 2 3; do
  WSB_LIBRARY=$PWD/nvidia-jetson-workload_b200/lib/libweather_b200_r1.so $B --steps 100 --warmup 10 | line "rk4 r1     rep$rep"
  $B --steps 100 --warmup 10 | line "rk4 strict rep$rep"
  WSB_ARITH=folded $B --steps 100 --warmup 10 | line "rk4 folded rep$rep"
done
for w in swe8192_euler prim2048x64 baro16384_f64; do
  WSB_LIBRARY=$PWD/nvidia-jetson-workload_b200/lib/libweather_b200_r1.so $B --workload $w --steps 50 --warmup 10 | line "$w r1    "
  $B --workload $w --steps 50 --warmup 10 | line "$w strict"
  WSB_ARITH=folded $B --workload $w --steps 50 --warmup 10 | line "$w folded"
done
for mb in 12 14 18 20; do
  WSB_FUSED_ROWS_PER_CHUNK=64 $B --steps 50 --warmup 10 | line "rk4 strict rpc64 (minb fixed)" ; break
done
for rpc in 48 96 128; do
  WSB_FUSED_ROWS_PER_CHUNK=$rpc $B --steps 50 --warmup 10 | line "rk4 strict rpc$rpc"
  WSB_ARITH=folded WSB_FUSED_ROWS_PER_CHUNK=$rpc $B --steps 50 --warmup 10 | line "rk4 folded rpc$rpc"
done
} 2>&1 | tee gpurun_out/r2a_ab.txt
B2="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e"
$B2 > gpurun_out/r2a_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_tma -s 5 -c 1 -f -o gpurun_out/prof_r2a_rk4 $B2 > gpurun_out/r2a_ncu.log 2>&1
echo "ncu strict: $?"
WSB_ARITH=folded $B2 > gpurun_out/r2a_plain_f.log 2>&1 && \
WSB_ARITH=folded ncu --set full --clock-control none --import-source on -k regex:step_tma -s 5 -c 1 -f -o gpurun_out/prof_r2a_rk4_folded $B2 > gpurun_out/r2a_ncu_f.log 2>&1
echo "ncu folded: $?"
$B2 > gpurun_out/r2a_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2a.csv $B2 > gpurun_out/r2a_ncu_l.log 2>&1
echo "ncu launches: $?"
ls -la gpurun_out | tail -20
