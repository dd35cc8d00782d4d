#!/usr/bin/env bash
# Round-2 GPU session W (2 GPUs): extended Primitive model on row slabs (tracer ghost rows), the new case of
# tests/test_multigpu.py alone, plus the single-GPU tests of the same model.
set -u
mkdir -p gpurun_out
timeout 100 python profiles/tools/check_slab_ext_primitive.py 2 > gpurun_out/r2w_slab_ext_primitive.txt 2>&1; echo "slab check rc $?"; tail -12 gpurun_out/r2w_slab_ext_primitive.txt
timeout 100 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "extended_primitive or primitive_levels" --tb=short -p no:cacheprovider > gpurun_out/r2w_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2w_pytest.log; tail -4 gpurun_out/r2w_pytest.log
