#!/usr/bin/env bash
# Round-1 GPU session D: full parity suite, the default bench line + reference arm, ncu launch list and full
# capture of the default bench command (short K), traffic numbers for profiles/traffic.json.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
python bench.py --impl reference --steps 50 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref exit $?"
cat gpurun_out/bench_default.json gpurun_out/bench_reference.json
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_default.csv $B > gpurun_out/ncu_l.log 2>&1
$B > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_tma -s 5 -c 2 -o gpurun_out/prof_step_fused_tma $B > gpurun_out/ncu_f.log 2>&1
E="python bench.py --workload swe8192_euler --steps 20 --warmup 3 --no-cpu-baseline --no-e2e"
$E > gpurun_out/plain_e.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_tma -s 5 -c 2 -o gpurun_out/prof_euler_tma $E > gpurun_out/ncu_e.log 2>&1
ls -la gpurun_out | tail -20
