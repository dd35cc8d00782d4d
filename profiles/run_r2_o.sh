#!/usr/bin/env bash
# Round-2 GPU session O (2 GPUs): slab parity on the final tree after the packed-division change (touches every
# whole-step kernel's source), plus the 2-GPU weak line.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q -rs --tb=short -p no:cacheprovider > gpurun_out/r2o_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2o_pytest.log
tail -6 gpurun_out/r2o_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-e2e --no-other-configs --steps 20 --warmup 5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n2 weak %.4f ms/step %.1f G' % (d['ms_per_step'], d['value']/1e9))" | tee gpurun_out/r2o_n2.txt
