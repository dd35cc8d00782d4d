#!/usr/bin/env bash
# Round-2 GPU session Q (1 GPU): ncu captures of the final tree (strict + folded RK4 kernel, launch list) so that
# profiles/instruction_mix.json / traffic.json carry the stamps of the shipped sources; default bench + reference arm.
set -u
mkdir -p gpurun_out
B2="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs"
$B2 > gpurun_out/r2q_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_tma -s 5 -c 1 -f -o gpurun_out/prof_r2q_rk4 $B2 > gpurun_out/r2q_ncu.log 2>&1
echo "ncu strict: $?"
$B2 --arith folded > gpurun_out/r2q_plain_f.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_tma -s 5 -c 1 -f -o gpurun_out/prof_r2q_rk4_folded $B2 --arith folded > gpurun_out/r2q_ncu_f.log 2>&1
echo "ncu folded: $?"
$B2 > gpurun_out/r2q_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2q.csv $B2 > gpurun_out/r2q_ncu_l.log 2>&1
echo "ncu launches: $?"
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2q_bench_default.json 2> gpurun_out/r2q_bench_default.err; echo "default bench rc $?"; tail -3 gpurun_out/r2q_bench_default.err
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2q_bench_reference.json 2> gpurun_out/r2q_bench_reference.err; echo "reference rc $?"
