#!/usr/bin/env bash
# Round-2 GPU session I (1 GPU): the reference's own weather_simulation.py over the shim (the byte-compiled wrapper now
# travels as .pyc.bin), smoke with the opt-ins, default bench + reference arm of the final tree.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_reference_wrapper.py tests/test_shim_gpu.py tests/test_bench_gpu.py -m gpu -q -rs --tb=short -p no:cacheprovider > gpurun_out/r2i_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2i_pytest.log
tail -8 gpurun_out/r2i_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1; echo "smoke rc $?"; tail -8 gpurun_out/r2i_smoke.log
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2i_bench_default.json 2> gpurun_out/r2i_bench_default.err; echo "default bench rc $?"; tail -3 gpurun_out/r2i_bench_default.err
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2i_bench_reference.json 2> gpurun_out/r2i_bench_reference.err; echo "reference rc $?"
