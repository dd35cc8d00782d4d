#!/usr/bin/env bash
# Round-2 GPU session T (1 GPU): whole GPU suite, smoke and the default bench on the final tree (after the dead-flag removal).
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -rs --tb=short -p no:cacheprovider > gpurun_out/r2t_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2t_pytest.log
tail -6 gpurun_out/r2t_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2t_smoke.log 2>&1; echo "smoke rc $?"; tail -8 gpurun_out/r2t_smoke.log
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2t_bench_default.json 2> gpurun_out/r2t_bench_default.err; echo "default bench rc $?"; tail -3 gpurun_out/r2t_bench_default.err
