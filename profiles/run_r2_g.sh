#!/usr/bin/env bash
# Round-2 GPU session G (1 GPU): the whole parity suite on the current tree (exact three-operation division, device
# side initial conditions, reference wrapper over the shim), division A/B, initial-condition timing at 32768^2, the
# default bench line and reference arm, ncu capture + launch list of the shipped RK4 kernel.
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2g_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2g_pytest.log
tail -12 gpurun_out/r2g_pytest.log
B="python bench.py --no-cpu-baseline --no-e2e --no-other-configs"
line() { python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$1 %.4f ms/step %.2f Gcell/s frac %.3f clk %s' % (d['ms_per_step'], d['value']/1e9, d['roofline']['frac'], d['clocks']['sm_mhz']))"; }
{
for rep in 1; do
  WSB_IEEE_DIV=1 $B --workload swe8192_rk4_div --steps 30 --warmup 5 | line "swe8192_rk4_div IEEE division      rep$rep"
  $B --workload swe8192_rk4_div --steps 30 --warmup 5 | line "swe8192_rk4_div 3-operation division rep$rep"
done
$B --workload swe8192_rk4_div --kernel stage_direct --steps 20 --warmup 5 | line "swe8192_rk4_div stage_direct 3-op"
$B --steps 100 --warmup 10 | line "rk4 strict (shipped defaults)"
$B --arith folded --steps 100 --warmup 10 | line "rk4 folded (shipped defaults)"
} 2>&1 | tee gpurun_out/r2g_ab.txt
timeout 300 python profiles/tools/ic_timing.py 32768 2>&1 | tee gpurun_out/r2g_ic_timing.txt
WSB_IC_HOST=1 timeout 240 python profiles/tools/ic_timing.py 16384 2>&1 | tee gpurun_out/r2g_ic_timing_host.txt
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2g_bench_default.json 2> gpurun_out/r2g_bench_default.err; echo "default bench rc $?"; tail -3 gpurun_out/r2g_bench_default.err
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2g_bench_reference.json 2> gpurun_out/r2g_bench_reference.err; echo "reference rc $?"
B2="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs"
$B2 > gpurun_out/r2g_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_tma -s 5 -c 1 -f -o gpurun_out/prof_r2g_rk4 $B2 > gpurun_out/r2g_ncu.log 2>&1
echo "ncu strict: $?"
$B2 > gpurun_out/r2g_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2g.csv $B2 > gpurun_out/r2g_ncu_l.log 2>&1
echo "ncu launches: $?"
