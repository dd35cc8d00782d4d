#!/usr/bin/env bash
# Round-2 GPU session T/X (1 GPU): whole GPU suite, smoke and the default bench on the final tree.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -rs --tb=short -p no:cacheprovider > gpurun_out/r2x_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2x_pytest.log
tail -6 gpurun_out/r2x_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2x_smoke.log 2>&1; echo "smoke rc $?"; tail -8 gpurun_out/r2x_smoke.log
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2x_bench_default.json 2> gpurun_out/r2x_bench_default.err; echo "default bench rc $?"; tail -3 gpurun_out/r2x_bench_default.err
