#!/usr/bin/env bash
# Round-1 GPU session: parity suite, tuning sweeps, ncu launch lists and full captures (run under gpurun).
set -u
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e"
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
{
for minb in 8 12 16; do for rpc in 64 128 256 512; do
  echo "== MINB=$minb RPC=$rpc"; WSB_FUSED_MINB=$minb WSB_FUSED_ROWS_PER_CHUNK=$rpc $B --steps 30 --warmup 5 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['clocks'])"
done; done
} > gpurun_out/sweep_fused.log 2>&1
cat gpurun_out/sweep_fused.log
# ncu: launch list then full capture of each kernel (the plain command first, && directly before ncu)
$B --steps 3 --warmup 2 > gpurun_out/plain_fused.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_fused.csv $B --steps 3 --warmup 2 > gpurun_out/ncu_l1.log 2>&1
$B --steps 3 --warmup 2 > gpurun_out/plain_fused.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_fused -s 3 -c 2 -o gpurun_out/prof_fused $B --steps 3 --warmup 2 > gpurun_out/ncu_f1.log 2>&1
$B --kernel stage_direct --steps 3 --warmup 2 > gpurun_out/plain_direct.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_direct.csv $B --kernel stage_direct --steps 3 --warmup 2 > gpurun_out/ncu_l2.log 2>&1
$B --kernel stage_direct --steps 3 --warmup 2 > gpurun_out/plain_direct.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:stage_direct -s 8 -c 4 -o gpurun_out/prof_direct $B --kernel stage_direct --steps 3 --warmup 2 > gpurun_out/ncu_f2.log 2>&1
ls -la gpurun_out
