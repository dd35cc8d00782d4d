#!/usr/bin/env bash
# Round-2 GPU session J (1 GPU): full parity suite of the final tree (extended Primitive tracers included), default
# bench + reference arm, final ncu captures (strict + folded) for profiles/instruction_mix.json / traffic.json, launch list.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rs --tb=short -p no:cacheprovider > gpurun_out/r2j_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2j_pytest.log
tail -12 gpurun_out/r2j_pytest.log
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2j_bench_default.json 2> gpurun_out/r2j_bench_default.err; echo "default bench rc $?"; tail -3 gpurun_out/r2j_bench_default.err
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2j_bench_reference.json 2> gpurun_out/r2j_bench_reference.err; echo "reference rc $?"
B2="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs"
$B2 > gpurun_out/r2j_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_tma -s 5 -c 1 -f -o gpurun_out/prof_r2j_rk4 $B2 > gpurun_out/r2j_ncu.log 2>&1
echo "ncu strict: $?"
$B2 --arith folded > gpurun_out/r2j_plain_f.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_tma -s 5 -c 1 -f -o gpurun_out/prof_r2j_rk4_folded $B2 --arith folded > gpurun_out/r2j_ncu_f.log 2>&1
echo "ncu folded: $?"
$B2 > gpurun_out/r2j_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2j.csv $B2 > gpurun_out/r2j_ncu_l.log 2>&1
echo "ncu launches: $?"
python bench.py --workload prim2048x64 --no-cpu-baseline --no-e2e --steps 20 --warmup 5 > gpurun_out/r2j_prim.json 2>/dev/null; echo "prim rc $?"
