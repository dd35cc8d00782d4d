#!/usr/bin/env bash
# Round-2 GPU session F (8 GPUs): slab parity at 2, 4 and 8 ranks on the current kernels (fused ghost exchange over
# peer memory), weak scaling of BASELINE config 2 at N = 1, 2, 4, 8 with the driver's own step counts (20 after 5) and
# with 200 steps, the NCCL path beside it, strong scaling of config 5 (sub-line of the bench), full N=8 bench line.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_multigpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2f_pytest.log
tail -15 gpurun_out/r2f_pytest.log
line() { python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 %.4f ms/step %.2f Gcell/s ranks %s' % (d['ms_per_step'], d['value']/1e9, [round(x,4) for x in d.get('ms_per_step_per_rank',[])]))"; }
{
python bench.py --no-cpu-baseline --no-e2e --no-other-configs --steps 20 --warmup 5 | line "n1 20 steps"
python bench.py --no-cpu-baseline --no-e2e --no-other-configs --steps 200 --warmup 10 | line "n1 200 steps"
for n in 2 4 8; do
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus $n --no-e2e --no-other-configs --steps 20 --warmup 5 2>/dev/null | line "n$n fused 20 steps"
$T bench.py --gpus $n --no-e2e --no-other-configs --steps 200 --warmup 10 2>/dev/null | line "n$n fused 200 steps"
WSB_NO_PEER_EXCHANGE=1 $T bench.py --gpus $n --no-e2e --no-other-configs --steps 200 --warmup 10 2>/dev/null | line "n$n NCCL path 200 steps"
done
} 2>&1 | tee gpurun_out/r2f_scale.txt
for n in 2 4 8; do
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2f_bench_n$n.json 2> gpurun_out/r2f_bench_n$n.err; echo "n$n full bench rc $?"
done
python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo "n1 full bench rc $?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --impl reference --steps 20 --warmup 5 > gpurun_out/r2f_bench_reference_n8.json 2> gpurun_out/r2f_bench_reference_n8.err; echo "reference arm n8 rc $?"
