#!/usr/bin/env bash
# Round-1 GPU session B: parity of all variants, sweep of the TMA-staged whole-step kernel, ncu captures.
set -u
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e"
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -12 gpurun_out/pytest.log
{
for k in step_fused_tma; do for rpc in 64 128 256; do
  echo "== $k RPC=$rpc"; WSB_FUSED_ROWS_PER_CHUNK=$rpc $B --kernel $k --steps 30 --warmup 5 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['clocks'])"
done; done
echo "== step_fused_reg MINB=12 RPC=64"; WSB_FUSED_MINB=12 WSB_FUSED_ROWS_PER_CHUNK=64 $B --kernel step_fused_reg --steps 30 --warmup 5 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['clocks'])"
} > gpurun_out/sweep_tma.log 2>&1
cat gpurun_out/sweep_tma.log
$B --kernel step_fused_tma --steps 3 --warmup 2 > gpurun_out/plain_tma.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_tma -s 3 -c 2 -o gpurun_out/prof_tma $B --kernel step_fused_tma --steps 3 --warmup 2 > gpurun_out/ncu_tma.log 2>&1
tail -3 gpurun_out/ncu_tma.log
ls -la gpurun_out | head -30
