#!/usr/bin/env bash
# Round-2 GPU session E (2 GPUs): fused ghost exchange over peer memory (one launch per step and rank, band CTAs
# wait / push / signal over NVLink, step overlap on slabs): parity, then weak scaling against the NCCL path
# (WSB_NO_PEER_EXCHANGE=1) and the single-GPU number.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2e_pytest.log
tail -15 gpurun_out/r2e_pytest.log
WSB_NO_PEER_EXCHANGE=1 timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2e_pytest_nccl.log 2>&1; echo "pytest (NCCL path) exit $?" >> gpurun_out/r2e_pytest_nccl.log
tail -3 gpurun_out/r2e_pytest_nccl.log
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
line() { python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 %.4f ms/step %.2f Gcell/s ranks %s' % (d['ms_per_step'], d['value']/1e9, d.get('ms_per_step_per_rank')))"; }
{
python bench.py --no-cpu-baseline --no-e2e --no-other-configs --steps 20 --warmup 5 | line "n1 20 steps"
python bench.py --no-cpu-baseline --no-e2e --no-other-configs --steps 200 --warmup 10 | line "n1 200 steps"
for rep in 1 2; do
$T bench.py --gpus 2 --no-e2e --no-other-configs --steps 20 --warmup 5 2>/dev/null | line "n2 fused 20 steps rep$rep"
$T bench.py --gpus 2 --no-e2e --no-other-configs --steps 200 --warmup 10 2>/dev/null | line "n2 fused 200 steps rep$rep"
WSB_NO_PEER_EXCHANGE=1 $T bench.py --gpus 2 --no-e2e --no-other-configs --steps 200 --warmup 10 2>/dev/null | line "n2 NCCL 200 steps rep$rep"
done
$T bench.py --gpus 2 --no-e2e --no-other-configs --arith folded --steps 200 --warmup 10 2>/dev/null | line "n2 fused folded 200 steps"
} 2>&1 | tee gpurun_out/r2e_scale.txt
$T bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err; echo "n2 full bench rc $?"; tail -3 gpurun_out/r2e_bench_n2.err
