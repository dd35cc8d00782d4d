#!/usr/bin/env bash
# Round-2 GPU session N (1 GPU): packed exact division for non-power-of-two spacing (per-row ordinary-input flags, scalar
# IEEE path for suspect rows): full parity suite, timing of the true-division configuration, default RK4 line.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/r2n_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2n_pytest.log
tail -6 gpurun_out/r2n_pytest.log
B="python bench.py --no-cpu-baseline --no-e2e --no-other-configs"
line() { python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$1 %.4f ms/step %.2f Gcell/s frac %.3f clk %s' % (d['ms_per_step'], d['value']/1e9, d['roofline']['frac'], d['clocks']['sm_mhz']))"; }
{
$B --workload swe8192_rk4_div --steps 30 --warmup 5 | line "swe8192_rk4_div packed 3-operation division rep1"
$B --workload swe8192_rk4_div --steps 30 --warmup 5 | line "swe8192_rk4_div packed 3-operation division rep2"
WSB_IEEE_DIV=1 $B --workload swe8192_rk4_div --steps 30 --warmup 5 | line "swe8192_rk4_div IEEE (no proven reciprocal: scalar path)"
$B --steps 100 --warmup 10 | line "rk4 strict"
$B --arith folded --steps 100 --warmup 10 | line "rk4 folded"
} 2>&1 | tee gpurun_out/r2n_ab.txt
