#!/usr/bin/env bash
# Round-2 GPU session K (8 GPUs): slab parity at 2, 4 and 8 ranks on the FINAL tree, then the 8-GPU weak-scaling line
# (driver protocol: 20 steps after 5) with the final bench.py.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -rs --tb=short -p no:cacheprovider > gpurun_out/r2k_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2k_pytest.log
tail -8 gpurun_out/r2k_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --no-e2e --no-other-configs --steps 20 --warmup 5 > gpurun_out/r2k_bench_n8_weak.json 2> gpurun_out/r2k_bench_n8_weak.err; echo "n8 weak rc $?"
python -c "
import json
d=json.loads(open('gpurun_out/r2k_bench_n8_weak.json').read().strip().splitlines()[-1]); print('n8 weak %.4f ms/step %.1f G' % (d['ms_per_step'], d['value']/1e9), d['halo']['exchange_us'])"
