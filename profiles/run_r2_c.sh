#!/usr/bin/env bash
# Round-2 GPU session C: step overlap with a provably convergent wait (no WARPSYNC.COLLECTIVE fallback code), A/B
# overlap on/off, rows-per-chunk sweep with overlap, the full parity suite, the default bench line + reference arm,
# ncu capture + launch list.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest.log
tail -4 gpurun_out/r2c_pytest.log
B="python bench.py --no-cpu-baseline --no-e2e --no-other-configs"
line() { python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$1 %.4f ms/step %.2f Gcell/s frac %.3f clk %s' % (d['ms_per_step'], d['value']/1e9, d['roofline']['frac'], d['clocks']['sm_mhz']))"; }
{
for rep in 1 2 3; do
  WSB_STEP_OVERLAP=0 $B --steps 100 --warmup 10 | line "rk4 strict nooverlap rep$rep"
  $B --steps 100 --warmup 10 | line "rk4 strict overlap   rep$rep"
  WSB_STEP_OVERLAP=0 $B --arith folded --steps 100 --warmup 10 | line "rk4 folded nooverlap rep$rep"
  $B --arith folded --steps 100 --warmup 10 | line "rk4 folded overlap   rep$rep"
done
for w in swe8192_euler baro16384_f64 swe8192_rk4_div; do for rep in 1 2; do
  WSB_STEP_OVERLAP=0 $B --workload $w --steps 50 --warmup 10 | line "$w nooverlap rep$rep"
  $B --workload $w --steps 50 --warmup 10 | line "$w overlap   rep$rep"
done; done
$B --workload prim2048x64 --steps 50 --warmup 10 | line "prim2048x64"
for rpc in 64 88 112 136 160 192 256; do
  WSB_FUSED_ROWS_PER_CHUNK=$rpc $B --steps 50 --warmup 10 | line "rk4 strict overlap rpc$rpc"
done
for rpc in 88 128 192; do
  WSB_FUSED_ROWS_PER_CHUNK=$rpc $B --arith folded --steps 50 --warmup 10 | line "rk4 folded overlap rpc$rpc"
  WSB_FUSED_ROWS_PER_CHUNK=$rpc $B --workload swe8192_euler --steps 50 --warmup 10 | line "euler overlap rpc$rpc"
  WSB_FUSED_ROWS_PER_CHUNK=$rpc $B --workload baro16384_f64 --steps 30 --warmup 5 | line "baro overlap rpc$rpc"
done
} 2>&1 | tee gpurun_out/r2c_ab.txt
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2c_bench_default.json 2> gpurun_out/r2c_bench_default.err; echo "default bench rc $?"; tail -3 gpurun_out/r2c_bench_default.err
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2c_bench_reference.json 2> gpurun_out/r2c_bench_reference.err; echo "reference rc $?"; tail -3 gpurun_out/r2c_bench_reference.err
B2="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs"
$B2 > gpurun_out/r2c_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_tma -s 5 -c 1 -f -o gpurun_out/prof_r2c_rk4 $B2 > gpurun_out/r2c_ncu.log 2>&1
echo "ncu strict: $?"
$B2 > gpurun_out/r2c_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2c.csv $B2 > gpurun_out/r2c_ncu_l.log 2>&1
echo "ncu launches: $?"
ls -la gpurun_out | tail -12
