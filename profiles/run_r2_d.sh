#!/usr/bin/env bash
# Round-2 GPU session D (2 GPUs): slab parity incl. multi-level Primitive and the pybind surface, weak-scaling bench
# with the device-side rendezvous, the bare ghost-exchange timing, and the no-exchange timing diagnostic
# (WSB_DEBUG_NO_EXCHANGE existed for this session only -- one launch per rank, no ghost exchange, wrong numbers next to
# the seams, timing only -- and was removed from the library afterwards).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2d_pytest.log
tail -5 gpurun_out/r2d_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
line() { python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 %.4f ms/step %.2f Gcell/s ranks %s halo %s' % (d['ms_per_step'], d['value']/1e9, d.get('ms_per_step_per_rank'), d.get('halo',{}).get('exchange_us')))"; }
{
python bench.py --no-cpu-baseline --no-e2e --no-other-configs --steps 20 --warmup 5 | line "n1 20 steps"
python bench.py --no-cpu-baseline --no-e2e --no-other-configs --steps 200 --warmup 10 | line "n1 200 steps"
for rep in 1 2; do
$T bench.py --gpus 2 --no-e2e --no-other-configs --steps 20 --warmup 5 2>/dev/null | line "n2 20 steps rep$rep"
$T bench.py --gpus 2 --no-e2e --no-other-configs --steps 200 --warmup 10 2>/dev/null | line "n2 200 steps rep$rep"
WSB_DEBUG_NO_EXCHANGE=1 $T bench.py --gpus 2 --no-e2e --no-other-configs --steps 200 --warmup 10 2>/dev/null | line "n2 200 steps NO EXCHANGE (timing only) rep$rep"
done
} 2>&1 | tee gpurun_out/r2d_scale.txt
$T bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err; echo "n2 full bench rc $?"; tail -3 gpurun_out/r2d_bench_n2.err
