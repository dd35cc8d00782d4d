#!/usr/bin/env bash
# Round-2 GPU session B: step overlap (programmatic launch + chunk-row counters, WSB_STEP_OVERLAP=0/1) and
# branch-free steady-state triple body A/B (lib/libweather_b200_nosteady.so = same sources
# with -DWSB_STEADY_BODY=0), parity of the changed kernel, the full default bench line (other_configs, strong sub-line,
# pybind leg) and the reference arm, ncu capture of the steady-body kernel.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_shim_gpu.py tests/test_bench_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_pytest.log
tail -4 gpurun_out/r2b_pytest.log
B="python bench.py --no-cpu-baseline --no-e2e --no-other-configs"
line() { python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$1 %.4f ms/step %.2f Gcell/s frac %.3f clk %s' % (d['ms_per_step'], d['value']/1e9, d['roofline']['frac'], d['clocks']['sm_mhz']))"; }
{
for rep in 1 2 3; do
  WSB_STEP_OVERLAP=0 WSB_LIBRARY=$PWD/nvidia-jetson-workload_b200/lib/libweather_b200_nosteady.so $B --steps 100 --warmup 10 | line "rk4 strict nosteady nooverlap rep$rep"
  WSB_STEP_OVERLAP=0 $B --steps 100 --warmup 10 | line "rk4 strict steady   nooverlap rep$rep"
  $B --steps 100 --warmup 10 | line "rk4 strict steady   overlap   rep$rep"
  WSB_STEP_OVERLAP=0 $B --arith folded --steps 100 --warmup 10 | line "rk4 folded steady   nooverlap rep$rep"
  $B --arith folded --steps 100 --warmup 10 | line "rk4 folded steady   overlap   rep$rep"
done
for w in swe8192_euler prim2048x64 baro16384_f64; do
  WSB_STEP_OVERLAP=0 $B --workload $w --steps 50 --warmup 10 | line "$w nooverlap"
  $B --workload $w --steps 50 --warmup 10 | line "$w overlap  "
done
for rpc in 40 52 64 76 88; do
  WSB_FUSED_ROWS_PER_CHUNK=$rpc $B --steps 50 --warmup 10 | line "rk4 strict steady rpc$rpc"
done
} 2>&1 | tee gpurun_out/r2b_ab.txt
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2b_bench_default.json 2> gpurun_out/r2b_bench_default.err; echo "default bench rc $?"; tail -3 gpurun_out/r2b_bench_default.err
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2b_bench_reference.json 2> gpurun_out/r2b_bench_reference.err; echo "reference rc $?"; tail -3 gpurun_out/r2b_bench_reference.err
B2="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs"
$B2 > gpurun_out/r2b_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_tma -s 5 -c 1 -f -o gpurun_out/prof_r2b_rk4 $B2 > gpurun_out/r2b_ncu.log 2>&1
echo "ncu strict: $?"
ls -la gpurun_out | tail -12
