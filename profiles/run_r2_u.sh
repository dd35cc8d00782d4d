#!/usr/bin/env bash
# Round-2 GPU session U (1 GPU): bench contract test with the extended cpu_baseline, default bench of the final tree,
# then the step-overlap soak (3000 steps at 8192^2: one enqueue == irregular batches == per-stage kernel, bit for bit).
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_bench_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2u_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2u_pytest.log
tail -4 gpurun_out/r2u_pytest.log
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2u_bench_default.json 2> gpurun_out/r2u_bench_default.err; echo "default bench rc $?"; tail -3 gpurun_out/r2u_bench_default.err
timeout 120 python profiles/tools/soak.py 3000 > gpurun_out/r2u_soak.txt 2>&1; echo "soak rc $?"; cat gpurun_out/r2u_soak.txt
