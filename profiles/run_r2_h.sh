#!/usr/bin/env bash
# Round-2 GPU session H (1 GPU): full parity suite with skip reasons on the final tree (out-of-line IEEE slow path of
# the exact division), division timing, default bench line.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rs --tb=short -p no:cacheprovider > gpurun_out/r2h_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2h_pytest.log
tail -12 gpurun_out/r2h_pytest.log
B="python bench.py --no-cpu-baseline --no-e2e --no-other-configs"
line() { python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$1 %.4f ms/step %.2f Gcell/s frac %.3f clk %s' % (d['ms_per_step'], d['value']/1e9, d['roofline']['frac'], d['clocks']['sm_mhz']))"; }
{
WSB_IEEE_DIV=1 $B --workload swe8192_rk4_div --steps 30 --warmup 5 | line "swe8192_rk4_div IEEE division (out of line)"
$B --workload swe8192_rk4_div --steps 30 --warmup 5 | line "swe8192_rk4_div 3-operation division"
$B --workload swe8192_rk4_div --kernel stage_direct --steps 20 --warmup 5 | line "swe8192_rk4_div stage_direct 3-op"
$B --workload swe8192_rk4_ext --steps 30 --warmup 5 | line "swe8192_rk4_ext (classical RK4, extended physics)"
$B --workload swe8192_rk4_ext --kernel stage_direct --steps 20 --warmup 5 | line "swe8192_rk4_ext stage_direct"
$B --steps 100 --warmup 10 | line "rk4 strict"
} 2>&1 | tee gpurun_out/r2h_ab.txt
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2h_bench_default.json 2> gpurun_out/r2h_bench_default.err; echo "default bench rc $?"; tail -3 gpurun_out/r2h_bench_default.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; echo "smoke rc $?"; tail -6 gpurun_out/r2h_smoke.log
